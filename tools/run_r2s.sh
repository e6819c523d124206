#!/bin/bash
# feature-head kernels: parity against PyTorch's operators, a stand-alone timing at the c3 shape, then the bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python -m pytest tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -4
timeout -s KILL 120 python - <<'PY'
import sys, torch
sys.path.insert(0, 'feature-3dgs_b200')
from diff_gaussian_rasterization import feature_head as fh
C, H, W = 128, 822, 1237
x = torch.randn(C, H, W, device='cuda'); gt = torch.randn(C, int(H / 2.25), int(W / 2.25), device='cuda')
for _ in range(3): fh.feature_l1_loss_and_grad(x, gt, 1.0)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20): fh.feature_l1_loss_and_grad(x, gt, 1.0)
e[1].record(); torch.cuda.synchronize()
print('feature head fwd+bwd ms', e[0].elapsed_time(e[1]) / 20)
PY
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2s_bench_c3.json 2> $O/r2s_bench_c3.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2s_bench_c3.json'))
print('value',b['value'],'ms/step',b['ms_per_step'],'e2e',b['e2e'],'launches',b['gpu_launches'],'other',b['config']['other_api'])
PY
tail -2 $O/r2s_bench_c3.err
