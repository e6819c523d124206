#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python tools/variant_times.py c3 8 nopf base lc16 lc16nopf > $O/r2q_c3.jsonl 2> $O/r2q_c3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2q_c3.jsonl'):
    d=json.loads(l); print(d['variant'], 'comp_bwd', d['ms']['comp_bwd'], 'g_feature diff', d.get('max_diff_vs_nopf',{}).get('g_feature'))
PY
tail -2 $O/r2q_c3.err
