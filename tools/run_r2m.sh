#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
el "feature head tests"
timeout -s KILL 300 python -m pytest tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -3
el "bench c3 (e2e with the fused head)"
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2m_bench_c3.json 2> $O/r2m_bench_c3.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2m_bench_c3.json'))
print('value',b['value'],'e2e',b['e2e'],'other',b['config']['other_api'])
PY
tail -2 $O/r2m_bench_c3.err
el "c5 / c2 forward: tensor-core path for C >= 16 vs fp32 path"
timeout -s KILL 300 python tools/variant_times.py c5 3 base > $O/r2m_c5_fp32.jsonl 2> $O/r2m_c5.err; cut -c1-300 $O/r2m_c5_fp32.jsonl
F3DGS_TC_MIN_C=16 timeout -s KILL 300 python tools/variant_times.py c5 3 base > $O/r2m_c5_tc.jsonl 2>> $O/r2m_c5.err; cut -c1-300 $O/r2m_c5_tc.jsonl
F3DGS_TC_MIN_C=16 timeout -s KILL 300 python tools/variant_times.py c2 5 base > $O/r2m_c2_tc.jsonl 2>> $O/r2m_c5.err; cut -c1-300 $O/r2m_c2_tc.jsonl
F3DGS_TC_MIN_C=16 timeout -s KILL 300 python tools/variant_times.py c3_C64 5 base > $O/r2m_c3c64_tc.jsonl 2>> $O/r2m_c5.err; cut -c1-300 $O/r2m_c3c64_tc.jsonl
timeout -s KILL 300 python tools/variant_times.py c3_C64 5 base > $O/r2m_c3c64_fp32.jsonl 2>> $O/r2m_c5.err; cut -c1-300 $O/r2m_c3c64_fp32.jsonl
tail -3 $O/r2m_c5.err
el "reference c2 again (twice)"
for i in 1 2; do timeout -s KILL 200 python bench.py --impl reference --config c2 --steps 5 --warmup 3 --l2-flush --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('ref c2', b['value'], b['ms_per_step'], b['e2e']['value'])"; done
timeout -s KILL 200 python bench.py --impl reference --config c2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('ref c2 noflush', b['value'], b['ms_per_step'], b['e2e']['value'])"
el done
