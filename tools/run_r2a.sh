#!/bin/bash
# First GPU call of round 2 (prepared at the end of round 1, when the GPU budget was spent):
#   1. every variant library timed and cross-checked in ONE python process (tools/variant_times.py, ~1 s per variant)
#   2. ncu --set full of the two composite kernels of the default library + raw CSV
#   3. ncu launch list of the bench command
# Run tools/prep_r2a.sh on the CPU side first.  Each step has its own timeout and writes into gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
# A fresh box pages python/torch/CUDA libraries in from cold storage: the first import can take minutes (round 1 lost four
# 120-second steps to this).  Pay it once, untimed, before anything with a tight time-out.
el "warm-up import"
timeout -s KILL 420 python -c "import torch, numpy; torch.zeros(8, device='cuda').sum().item(); print('torch', torch.__version__, torch.cuda.get_device_name(0))"
el variants
timeout -s KILL 400 python tools/variant_times.py c3 5 base u0f0 u1f0 u1f1p ec pr ecpr nofr nogr nofma timing > $O/r2a_variants.jsonl 2> $O/r2a_variants.err
cat $O/r2a_variants.jsonl | cut -c1-400; grep "f3dgs timing" $O/r2a_variants.err | tail -2
el "risky variants (own process: a hang must not take the sweep down)"
timeout -s KILL 150 python tools/variant_times.py c3 5 base hp hpec > $O/r2a_variants_hp.jsonl 2>> $O/r2a_variants.err; cat $O/r2a_variants_hp.jsonl | cut -c1-400
el "two-pass mode (own process)"
F3DGS_SPLIT=1 timeout -s KILL 150 python tools/check_lists.py small 40 2>&1 | tail -2 | tee $O/r2a_lists.txt
timeout -s KILL 150 python tools/variant_times.py small 3 base base+split base+split2 > $O/r2a_split_small.jsonl 2>> $O/r2a_variants.err; cat $O/r2a_split_small.jsonl | cut -c1-600
timeout -s KILL 200 python tools/variant_times.py c3 5 base base+split base+split2 > $O/r2a_split_c3.jsonl 2>> $O/r2a_variants.err; cat $O/r2a_split_c3.jsonl | cut -c1-600
F3DGS_SPLIT=1 timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small_configs or feature_widths or image_shapes or c2_vs" 2>&1 | tail -3 | tee $O/r2a_split_pytest.txt
el "variants, config 2"
timeout -s KILL 200 python tools/variant_times.py c2 5 base u0f0 ec pr > $O/r2a_variants_c2.jsonl 2>> $O/r2a_variants.err; cat $O/r2a_variants_c2.jsonl | cut -c1-300
el "ncu full"
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:composite -s 2 -c 2 -f -o $O/prof_r2a_c3 python tools/prof_one.py c3 2 > $O/r2a_ncu_full.log 2>&1; tail -2 $O/r2a_ncu_full.log
timeout -s KILL 120 ncu -i $O/prof_r2a_c3.ncu-rep --page raw --csv > $O/prof_r2a_c3_raw.csv 2>/dev/null
el "ncu launch list"
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2a_launches.log 2>&1; tail -1 $O/r2a_launches.log | cut -c1-300
el done
