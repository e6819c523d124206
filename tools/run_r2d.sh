#!/bin/bash
# Round 2: tensor-core forward bring-up (F3DGS_TC=1) + the new view-batch / C-ABI tests.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
el "warm-up import"
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
el "tc forward: c3_C128 widths on the small / c3 scenes through the C ABI (one process, max diff vs the fp32-pipe kernels)"
timeout -s KILL 120 python tools/variant_times.py small128 3 base base+tc > $O/r2d_tc_small.jsonl 2> $O/r2d_tc_small.err; cut -c1-500 $O/r2d_tc_small.jsonl; tail -3 $O/r2d_tc_small.err
timeout -s KILL 120 python tools/variant_times.py small200 3 base base+tc > $O/r2d_tc_small200.jsonl 2> $O/r2d_tc_small200.err; cut -c1-500 $O/r2d_tc_small200.jsonl; tail -3 $O/r2d_tc_small200.err
timeout -s KILL 180 python tools/variant_times.py c3 5 base base+tc > $O/r2d_tc_c3.jsonl 2> $O/r2d_tc_c3.err; cut -c1-500 $O/r2d_tc_c3.jsonl; tail -3 $O/r2d_tc_c3.err
timeout -s KILL 180 python tools/variant_times.py c4 3 base base+tc > $O/r2d_tc_c4.jsonl 2> $O/r2d_tc_c4.err; cut -c1-500 $O/r2d_tc_c4.jsonl; tail -3 $O/r2d_tc_c4.err
el "pytest -m gpu with F3DGS_TC=1"
F3DGS_TC=1 timeout -s KILL 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^  " | tail -15 | tee $O/r2d_pytest_tc.txt
el "pytest -m gpu default (new tests)"
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "view_batch or c_abi" 2>&1 | tail -5 | tee $O/r2d_pytest_new.txt
el done
