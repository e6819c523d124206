#!/bin/bash
# compute-sanitizer memcheck of one small forward + backward (tensor-core feature-gradient kernel on) + feature head
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
F3DGS_FBWD_TC=1 timeout -s KILL 70 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_probe.py small128 > $O/r2v_memcheck.txt 2>&1
echo rc=$?; tail -6 $O/r2v_memcheck.txt
