#!/bin/bash
# Batch / handle sweep of the device-resident leg (one line per configuration).  usage: scripts/sweep.sh WORKLOAD "B H" "B H" ...
wl=${1:-sr_lo_lm}; shift
for cfg in "$@"; do
  set -- $cfg
  echo "== $wl batch $1 handles $2"
  timeout 500 python bench.py --workload $wl --batch $1 --handles $2 --legs device --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']))
print('  '+' '.join('%s=%.0f'%(k,v['avg_us']) for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['avg_us']*kv[1]['launches'])[:12]))"
done
