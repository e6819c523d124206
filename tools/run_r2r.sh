#!/bin/bash
# Round-2 check of the tensor-core feature-gradient kernel (F3DGS_FBWD_TC=1) against the fp32 one: parity on a small
# scene with ragged image sizes, then timing and role counters at c3.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 420 python -c "import torch; torch.zeros(8, device='cuda').sum().item()"
run() {
  cfg=$1; it=$2; shift 2
  timeout -s KILL 300 python tools/variant_times.py $cfg $it "$@" > $O/r2r_$cfg.jsonl 2> $O/r2r_$cfg.err
  echo "== $cfg rc=$?"
  python - $cfg <<'PY'
import json, sys
for l in open(f'gpurun_out/r2r_{sys.argv[1]}.jsonl'):
    d = json.loads(l)
    df = d.get('max_diff_vs_base') or {}
    print(d['variant'], 'comp_bwd', d['ms']['comp_bwd'], 'g_feature diff', df.get('g_feature'), 'max other', max([v for k, v in df.items() if k != 'g_feature'] or [0]))
PY
  grep fbtc_diag $O/r2r_$cfg.err | awk 'NR%5==3'
  grep -v fbtc_diag $O/r2r_$cfg.err | tail -2
}
run small200 2 base base+fbtc
run c3 8 base base+fbtc base+fbtc+nopf base+fbtc+diag base+fbtc+nopf+diag
