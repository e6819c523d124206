#!/bin/bash
# Round-2 check of the tensor-core feature-gradient kernel (F3DGS_FBWD_TC=1) against the fp32 one: parity on two small
# scenes (ragged image sizes), then timing at c3.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
for cfg in small200 c3; do
  it=8; [ $cfg != c3 ] && it=2
  timeout -s KILL 300 python tools/variant_times.py $cfg $it base base+fbtc base+fbtc+nohelp > $O/r2r_$cfg.jsonl 2> $O/r2r_$cfg.err
  echo "== $cfg rc=$?"
  python - $cfg <<'PY'
import json, sys
for l in open(f'gpurun_out/r2r_{sys.argv[1]}.jsonl'):
    d = json.loads(l)
    print(d['variant'], 'comp_bwd', d['ms']['comp_bwd'], 'diff', d.get('max_diff_vs_base'))
PY
  tail -2 $O/r2r_$cfg.err
done
