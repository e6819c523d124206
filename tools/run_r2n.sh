#!/bin/bash
# multi-GPU: NCCL correctness of ViewBatch, then bench at N GPUs (c3 weak, c4 strong)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 420 python -c "import torch; print(torch.cuda.device_count()); torch.zeros(8, device='cuda').sum().item()"
el "nccl check"
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_viewbatch_nccl.py small128 2 2>&1 | grep -E "viewbatch|Error|error" | head -5
el "bench c3 N=$N"
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > $O/r2n_c3_n$N.json 2> $O/r2n_c3_n$N.err; cut -c1-220 $O/r2n_c3_n$N.json; tail -2 $O/r2n_c3_n$N.err | cut -c1-300
el "bench c4 N=$N"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c4 --steps 3 --warmup 3 > $O/r2n_c4_n$N.json 2> $O/r2n_c4_n$N.err; cut -c1-220 $O/r2n_c4_n$N.json; tail -2 $O/r2n_c4_n$N.err | cut -c1-300
el done
