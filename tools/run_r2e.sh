#!/bin/bash
# Round 2: new defaults (tensor-core forward, two-kernel backward, exact cull, pair skip) -- parity + timing.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
el "variants c3 / c4 / c2 / small128"
timeout -s KILL 200 python tools/variant_times.py c3 5 base+notc+bwd1 base base+bwd1 base+notc > $O/r2e_c3.jsonl 2> $O/r2e_c3.err; cut -c1-520 $O/r2e_c3.jsonl; tail -3 $O/r2e_c3.err
timeout -s KILL 200 python tools/variant_times.py c4 3 base+notc+bwd1 base > $O/r2e_c4.jsonl 2> $O/r2e_c4.err; cut -c1-520 $O/r2e_c4.jsonl; tail -3 $O/r2e_c4.err
timeout -s KILL 200 python tools/variant_times.py c2 5 base+notc+bwd1 base > $O/r2e_c2.jsonl 2> $O/r2e_c2.err; cut -c1-520 $O/r2e_c2.jsonl; tail -3 $O/r2e_c2.err
el "pytest -m gpu (defaults)"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^  " | tail -12 | tee $O/r2e_pytest.txt
el "pytest -m gpu subset with the fused / fp32 fallbacks"
F3DGS_TC=0 F3DGS_BWD2=0 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_configs or feature_widths or image_shapes or c3_full" 2>&1 | tail -3 | tee $O/r2e_pytest_fallbacks.txt
el "bench c3"
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > $O/r2e_bench_c3.json 2> $O/r2e_bench_c3.err; cut -c1-400 $O/r2e_bench_c3.json; tail -2 $O/r2e_bench_c3.err
el done
