mkdir -p gpurun_out
for v in base st5 st8; do echo "== $v"; tools/with_variant.sh $v timeout -s KILL 90 python tools/stage_times.py c3 5 2>&1 | tail -1; done
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
