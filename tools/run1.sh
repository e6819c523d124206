mkdir -p gpurun_out
timeout -s KILL 120 python tools/stress_determinism.py c2 60 2>&1 | tail -4
timeout -s KILL 120 python tools/stress_determinism.py small 200 2>&1 | tail -3
timeout -s KILL 120 python tools/stress_determinism.py c3_C32 10 2>&1 | tail -3
timeout -s KILL 120 python tools/stress_determinism.py c1 100 2>&1 | tail -3
F3DGS_BPA=2 timeout -s KILL 120 python tools/stress_determinism.py c2 40 2>&1 | tail -3
timeout -s KILL 90 python tools/stage_times.py c3 5 2>&1 | tail -1
