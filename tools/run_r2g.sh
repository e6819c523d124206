#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 400 python tools/variant_times.py c3 5 tc1 tc1x tc1c8 base tc2c16 tc1nomma tc2nomma tc1noalpha tc1nozero tc1noload tc1onlymma > $O/r2g_c3.jsonl 2> $O/r2g_c3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2g_c3.jsonl'):
    d=json.loads(l); print(d['variant'], 'comp_fwd', d['ms']['comp_fwd'], 'feature diff', d.get('max_diff_vs_tc1',{}).get('feature'))
PY
tail -3 $O/r2g_c3.err
