#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
timeout -s KILL 300 python tools/variant_times.py c3 5 base tc2 timingtc > $O/r2j_c3.jsonl 2> $O/r2j_c3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2j_c3.jsonl'):
    d=json.loads(l); print(d['variant'], 'comp_fwd', d['ms']['comp_fwd'])
PY
grep "f3dgs timing" $O/r2j_c3.err | tail -2
