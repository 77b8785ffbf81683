#!/bin/bash
# handles / batch sweep (device leg only)
out=gpurun_out/sweep_s2a.txt; : > $out
for cfg in "sr_lo 256 1" "sr_lo 256 2" "sr_lo 256 3" "sr_lo 256 4" "sr_lo 512 4" "sr_lo_lm 32 1" "sr_lo_lm 32 2" "sr_lo_lm 32 4" "sr_lo_lm 64 2" "sr_lo_lm 64 4" "sr_lo_lm 128 4"; do
  set -- $cfg
  echo "== $cfg" >> $out
  timeout 300 python bench.py --workload $1 --batch $2 --handles $3 --legs device --steps 20 2>>gpurun_out/sweep_s2a.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']))
print('  '+' '.join('%s=%.0f'%(k,v['avg_us']) for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['avg_us']*kv[1]['launches'])[:10]))
" >> $out 2>&1
done
cat $out
