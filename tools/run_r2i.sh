#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python tools/variant_times.py c3 5 base tc1skel tc1noam > $O/r2i_c3.jsonl 2> $O/r2i_c3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2i_c3.jsonl'):
    d=json.loads(l); print(d['variant'], 'comp_fwd', d['ms']['comp_fwd'])
PY
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:composite_fwd_tc -s 1 -c 1 -f -o $O/prof_r2i_fwdtc python tools/prof_one.py c3 2 > $O/r2i_ncu.log 2>&1; tail -2 $O/r2i_ncu.log
timeout -s KILL 120 ncu -i $O/prof_r2i_fwdtc.ncu-rep --page source --csv > $O/prof_r2i_fwdtc_src.csv 2>/dev/null
timeout -s KILL 120 ncu -i $O/prof_r2i_fwdtc.ncu-rep --page raw --csv > $O/prof_r2i_fwdtc_raw.csv 2>/dev/null
ls -la $O/prof_r2i*
