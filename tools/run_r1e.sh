#!/bin/bash
# One GPU call (round 1, session e): micro-benchmarks, variant timings, full GPU suite on the default library,
# bench line, ncu full capture of the two composite kernels, diagnostics, ncu launch list of the bench command.
# Every step has its own timeout and writes into gpurun_out/ as it goes.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $O/r1e_gpu.txt 2>&1
el micro
( timeout -s KILL 30 tools/micro/red_rate; timeout -s KILL 20 tools/micro/ffma_rate; timeout -s KILL 20 tools/micro/ffma2_rate ) > $O/r1e_micro.txt 2>&1
cat $O/r1e_micro.txt
el variants
for v in base u0f0 u1f0 u1f1p; do
  echo "== $v"; tools/with_variant.sh $v timeout -s KILL 120 python tools/stage_times.py c3 4 2>&1 | tail -1
done | tee $O/r1e_variants.txt
el "full suite (default library)"
timeout -s KILL 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $O/r1e_pytest.txt
el bench
timeout -s KILL 300 python bench.py > $O/r1e_bench.json 2> $O/r1e_bench.err; tail -c 1200 $O/r1e_bench.json
el "ncu full"
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:composite -s 2 -c 2 -f -o $O/prof_r1e_c3 python tools/prof_one.py c3 2 > $O/r1e_ncu_full.log 2>&1; tail -2 $O/r1e_ncu_full.log
timeout -s KILL 120 ncu -i $O/prof_r1e_c3.ncu-rep --page raw --csv > $O/prof_r1e_c3_raw.csv 2>/dev/null
el diagnostics
( echo "== timing"; F3DGS_TIMING=1 tools/with_variant.sh timing timeout -s KILL 120 python tools/stage_times.py c3 2 2>&1 | grep "f3dgs timing" | tail -2
  for v in nofr nogr; do echo "== $v"; tools/with_variant.sh $v timeout -s KILL 120 python tools/stage_times.py c3 4 2>&1 | tail -1; done ) | tee $O/r1e_diag.txt
el "subset parity on variants"
for v in u1f0 u1f1p; do
  echo "== $v"; tools/with_variant.sh $v timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small_configs" 2>&1 | tail -2
done | tee $O/r1e_variant_parity.txt
el "ncu launch list"
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r1e_launches.log 2>&1; tail -1 $O/r1e_launches.log | cut -c1-300
el done
