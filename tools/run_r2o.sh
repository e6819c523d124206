#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python -m pytest tests/test_next_rows.py tests/test_gpu_parity.py -m gpu -x -q -k "feature_head or feature_widths or c3_full" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2o_bench_c3.json 2> $O/r2o_bench_c3.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2o_bench_c3.json'))
print('value',b['value'],'e2e',b['e2e'],'other',b['config']['other_api'])
PY
tail -2 $O/r2o_bench_c3.err
timeout -s KILL 300 python bench.py --config c5 --steps 5 --warmup 3 > $O/r2o_bench_c5.json 2> $O/r2o_bench_c5.err; cut -c1-200 $O/r2o_bench_c5.json
timeout -s KILL 200 python bench.py --impl reference --config c2 --steps 5 --warmup 3 --l2-flush --no-cpu-baseline > $O/r2o_ref_c2.json 2>/dev/null; cut -c1-200 $O/r2o_ref_c2.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize" -c 12 --csv --log-file $O/r2o_head_launches.csv python - <<'PY' > /dev/null 2>&1
import sys; sys.path.insert(0,'feature-3dgs_b200')
import torch
from diff_gaussian_rasterization import feature_head as fh
fm=torch.randn(128,1080,1920,device='cuda'); gt=torch.rand(128,480,853,device='cuda')
for _ in range(3):
    l,g=fh.feature_l1_loss_and_grad(fm,gt,1.0)
torch.cuda.synchronize()
PY
grep -E "resize" $O/r2o_head_launches.csv | cut -d, -f5,12- | cut -c1-200 | tail -12
