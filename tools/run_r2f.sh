#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
el "tc forward 2 CTAs/SM (base) vs 1 (tc1); backward geometry kernel 3 CTAs/SM (slim3: 4 red slots, slim3b: 5)"
timeout -s KILL 120 python tools/variant_times.py small128 3 tc1 base > $O/r2f_small.jsonl 2> $O/r2f_small.err; cut -c1-420 $O/r2f_small.jsonl; tail -3 $O/r2f_small.err
timeout -s KILL 240 python tools/variant_times.py c3 5 tc1 base slim3 slim3b > $O/r2f_c3.jsonl 2> $O/r2f_c3.err; cut -c1-420 $O/r2f_c3.jsonl; tail -3 $O/r2f_c3.err
timeout -s KILL 240 python tools/variant_times.py c4 3 tc1 base slim3b > $O/r2f_c4.jsonl 2> $O/r2f_c4.err; cut -c1-420 $O/r2f_c4.jsonl; tail -3 $O/r2f_c4.err
el "pytest subset"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_configs or feature_widths or c3_full or c4_full" 2>&1 | tail -3
el done
