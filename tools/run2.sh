mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or small" 2>&1 | tail -2
for b in 1 2; do F3DGS_BPA=$b timeout -s KILL 120 python tools/stage_times.py c3 5 2>&1 | tail -1; done
F3DGS_BPA=1 timeout -s KILL 280 ncu --set full --import-source on --clock-control none -k regex:composite_bwd -s 1 -c 1 -f -o gpurun_out/prof_r1d_bwd_bpa1 python tools/prof_one.py c3 2 > gpurun_out/ncu_e1.log 2>&1
tail -2 gpurun_out/ncu_e1.log
