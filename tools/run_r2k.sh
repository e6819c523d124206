#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python tools/variant_times.py c3 5 base > $O/r2k_c3.jsonl 2> $O/r2k_c3.err; cut -c1-400 $O/r2k_c3.jsonl; tail -2 $O/r2k_c3.err
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2k_pytest.txt
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > $O/r2k_bench_c3.json 2> $O/r2k_bench_c3.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2k_bench_c3.json'))
print('value',b['value'],'e2e',b['e2e']['value'],'other',b['config']['other_api'])
print(b['roofline']['stage_ms_per_launch'])
PY
tail -2 $O/r2k_bench_c3.err
