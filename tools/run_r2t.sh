#!/bin/bash
# final check of round 2: full GPU suite, smoke, default bench line (ours + a short reference-arm run), c2 line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2t_pytest.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > $O/r2t_bench_c3.json 2> $O/r2t_bench_c3.err
timeout -s KILL 200 python bench.py --config c2 --l2-flush --steps 10 --warmup 3 --no-cpu-baseline > $O/r2t_bench_c2.json 2> $O/r2t_bench_c2.err
python - <<'PY'
import json
for c in ('c3', 'c2'):
    b=json.load(open(f'gpurun_out/r2t_bench_{c}.json'))
    print(c, 'value',b['value'],'ms/step',b['ms_per_step'],'e2e',b['e2e'],'launches',b['gpu_launches'],'other',b['config']['other_api'], 'clocks', b.get('clocks'))
    print(b['roofline']['stage_ms_per_launch'], b['roofline']['frac'], b['roofline']['traffic'], b.get('cpu_baseline'))
PY
tail -2 $O/r2t_bench_c3.err
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 3 --no-cpu-baseline > $O/r2t_ref_c3.json 2> $O/r2t_ref_c3.err; cut -c1-300 $O/r2t_ref_c3.json; tail -2 $O/r2t_ref_c3.err
