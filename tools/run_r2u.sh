#!/bin/bash
# ncu --set full of the feature-head kernels at the config-3 shape
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 100 ncu --set full --clock-control none -k regex:resize_bwd -c 1 -f -o $O/r02_head_bwd python tools/head_probe.py 2>&1 | tail -5
ls -la $O/r02_head_bwd.ncu-rep
