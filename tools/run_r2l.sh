#!/bin/bash
# Round 2 evidence run (1 GPU): bench lines of every BASELINE config for both arms, ncu --set full + launch list of the
# default kernels.  Outputs are copied into profiles/ by the caller.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
el "bench ours"
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > $O/r2l_bench_c3.json 2> $O/r2l_bench_c3.err; cut -c1-250 $O/r2l_bench_c3.json; tail -2 $O/r2l_bench_c3.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --l2-flush --no-cpu-baseline > $O/r2l_bench_c3_flush.json 2>> $O/r2l_bench_c3.err; cut -c1-200 $O/r2l_bench_c3_flush.json
timeout -s KILL 200 python bench.py --config c2 --steps 20 --warmup 3 --l2-flush > $O/r2l_bench_c2.json 2> $O/r2l_bench_c2.err; cut -c1-250 $O/r2l_bench_c2.json; tail -2 $O/r2l_bench_c2.err
timeout -s KILL 400 python bench.py --config c4 --steps 3 --warmup 3 > $O/r2l_bench_c4.json 2> $O/r2l_bench_c4.err; cut -c1-250 $O/r2l_bench_c4.json; tail -2 $O/r2l_bench_c4.err
timeout -s KILL 300 python bench.py --config c5 --steps 5 --warmup 3 > $O/r2l_bench_c5.json 2> $O/r2l_bench_c5.err; cut -c1-250 $O/r2l_bench_c5.json; tail -2 $O/r2l_bench_c5.err
el "bench reference arm"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 3 --no-cpu-baseline > $O/r2l_ref_c3.json 2> $O/r2l_ref_c3.err; cut -c1-250 $O/r2l_ref_c3.json; tail -2 $O/r2l_ref_c3.err
timeout -s KILL 200 python bench.py --impl reference --config c2 --steps 5 --warmup 3 --l2-flush --no-cpu-baseline > $O/r2l_ref_c2.json 2> $O/r2l_ref_c2.err; cut -c1-250 $O/r2l_ref_c2.json
timeout -s KILL 400 python bench.py --impl reference --config c4 --batch-views 4 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2l_ref_c4.json 2> $O/r2l_ref_c4.err; cut -c1-250 $O/r2l_ref_c4.json
timeout -s KILL 300 python bench.py --impl reference --config c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2l_ref_c5.json 2> $O/r2l_ref_c5.err; cut -c1-250 $O/r2l_ref_c5.json
timeout -s KILL 300 python bench.py --impl reference --ref-debug --steps 2 --warmup 3 --no-cpu-baseline > $O/r2l_ref_c3_debug.json 2> $O/r2l_ref_c3_debug.err; cut -c1-250 $O/r2l_ref_c3_debug.json
el "ncu full (default kernels, c3)"
timeout -s KILL 500 ncu --set full --import-source on --clock-control none -k regex:"composite|feature_bwd|preprocess" -s 5 -c 5 -f -o $O/prof_r2l_c3 python tools/prof_one.py c3 2 > $O/r2l_ncu_full.log 2>&1; tail -2 $O/r2l_ncu_full.log
timeout -s KILL 120 ncu -i $O/prof_r2l_c3.ncu-rep --page raw --csv > $O/prof_r2l_c3_raw.csv 2>/dev/null
el "ncu launch list (ours, bench command)"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2l_launches.log 2>&1; tail -1 $O/r2l_launches.log | cut -c1-200
el done
