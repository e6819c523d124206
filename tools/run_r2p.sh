#!/bin/bash
# final check of a round: full GPU suite, smoke, default bench line (ours + a short reference-arm run)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2p_pytest.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > $O/r2p_bench_c3.json 2> $O/r2p_bench_c3.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2p_bench_c3.json'))
print('value',b['value'],'ms/step',b['ms_per_step'],'e2e',b['e2e'],'launches',b['gpu_launches'],'other',b['config']['other_api'])
print(b['roofline']['stage_ms_per_launch'], b['roofline']['frac'], b['roofline']['traffic'])
PY
tail -2 $O/r2p_bench_c3.err
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 3 --no-cpu-baseline > $O/r2p_ref_c3.json 2> $O/r2p_ref_c3.err; cut -c1-300 $O/r2p_ref_c3.json; tail -2 $O/r2p_ref_c3.err
