#!/bin/bash
# Round 2, GPU call #1: full GPU test-suite (incl. the new c4/c5/C-ABI tests), the prepared variant sweep, bench lines for
# c2/c3/c4/c5 (both arms), ncu --set full + launch list of the DEFAULT kernels.  CPU side first: tools/prep_r2a.sh.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
el "warm-up import"
timeout -s KILL 420 python -c "import torch, numpy; torch.zeros(8, device='cuda').sum().item(); print('torch', torch.__version__, torch.cuda.get_device_name(0))"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $O/r2b_smi.txt
el "tcgen05 tf32 probe"
timeout -s KILL 90 tools/micro/tc_probe 2>&1 | tee $O/r2b_tc_probe.txt
el "pytest -m gpu"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^  " | tail -25 | tee $O/r2b_pytest.txt
el variants
timeout -s KILL 400 python tools/variant_times.py c3 5 base u1f1p ec pr ecpr nofr nogr nofma timing > $O/r2b_variants.jsonl 2> $O/r2b_variants.err
cut -c1-420 $O/r2b_variants.jsonl; grep "f3dgs timing" $O/r2b_variants.err | tail -2
timeout -s KILL 150 python tools/variant_times.py c3 5 base hp hpec > $O/r2b_variants_hp.jsonl 2>> $O/r2b_variants.err; cut -c1-420 $O/r2b_variants_hp.jsonl
el "two-pass mode"
F3DGS_SPLIT=1 timeout -s KILL 150 python tools/check_lists.py small 40 2>&1 | tail -2 | tee $O/r2b_lists.txt
timeout -s KILL 200 python tools/variant_times.py c3 5 base base+split base+split2 > $O/r2b_split_c3.jsonl 2>> $O/r2b_variants.err; cut -c1-600 $O/r2b_split_c3.jsonl
el "bench ours c3 (default), c2, c4, c5"
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_c3.json 2> $O/r2b_bench_c3.err; cut -c1-300 $O/r2b_bench_c3.json
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --l2-flush --no-cpu-baseline > $O/r2b_bench_c3_flush.json 2>> $O/r2b_bench_c3.err; cut -c1-200 $O/r2b_bench_c3_flush.json
timeout -s KILL 200 python bench.py --config c2 --steps 20 --warmup 3 --l2-flush > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err; cut -c1-300 $O/r2b_bench_c2.json
timeout -s KILL 400 python bench.py --config c4 --steps 3 --warmup 3 > $O/r2b_bench_c4.json 2> $O/r2b_bench_c4.err; cut -c1-300 $O/r2b_bench_c4.json
timeout -s KILL 300 python bench.py --config c5 --steps 5 --warmup 3 > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; cut -c1-300 $O/r2b_bench_c5.json
el "bench reference arm c3, c2, c4 (4-view batch), c5; c3 with debug=True"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 3 --no-cpu-baseline > $O/r2b_ref_c3.json 2> $O/r2b_ref_c3.err; cut -c1-300 $O/r2b_ref_c3.json
timeout -s KILL 200 python bench.py --impl reference --config c2 --steps 5 --warmup 3 --l2-flush --no-cpu-baseline > $O/r2b_ref_c2.json 2> $O/r2b_ref_c2.err; cut -c1-300 $O/r2b_ref_c2.json
timeout -s KILL 400 python bench.py --impl reference --config c4 --batch-views 4 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_ref_c4.json 2> $O/r2b_ref_c4.err; cut -c1-300 $O/r2b_ref_c4.json
timeout -s KILL 300 python bench.py --impl reference --config c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2b_ref_c5.json 2> $O/r2b_ref_c5.err; cut -c1-300 $O/r2b_ref_c5.json
timeout -s KILL 300 python bench.py --impl reference --ref-debug --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_ref_c3_debug.json 2> $O/r2b_ref_c3_debug.err; cut -c1-300 $O/r2b_ref_c3_debug.json
el "ncu full (default kernels, c3)"
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:composite -s 2 -c 2 -f -o $O/prof_r2b_c3 python tools/prof_one.py c3 2 > $O/r2b_ncu_full.log 2>&1; tail -2 $O/r2b_ncu_full.log
timeout -s KILL 120 ncu -i $O/prof_r2b_c3.ncu-rep --page raw --csv > $O/prof_r2b_c3_raw.csv 2>/dev/null
el "ncu launch list (ours, bench command)"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2b_launches.log 2>&1; tail -1 $O/r2b_launches.log | cut -c1-200
el "ncu launch list (reference arm: one line for its backward renderCUDA)"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:renderCUDA -c 8 --csv --log-file $O/r2b_ref_launches.csv python bench.py --impl reference --steps 1 --warmup 1 --views-per-rank 1 --no-cpu-baseline > $O/r2b_ref_launches.log 2>&1; tail -3 $O/r2b_ref_launches.csv | cut -c1-300
el done
