#!/bin/bash
out=gpurun_out/sweep_s2l.txt; : > $out
python -m pytest tests/test_gpu_mapping.py -x -q 2>&1 | tail -2 >> $out
for cfg in "sr_lo_lm 128 2" "sr_lo_lm 128 4" "sr_lo_lm 192 3" "sr_lo_lm 256 2" "sr_lo_lm 256 4"; do
  set -- $cfg
  echo "== $cfg" >> $out
  timeout 500 python bench.py --workload $1 --batch $2 --handles $3 --legs device --steps 20 2>>gpurun_out/sweep_s2l.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']))
print('  '+' '.join('%s=%.0f'%(k,v['avg_us']) for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['avg_us']*kv[1]['launches'])[:12]))
" >> $out 2>&1
done
cat $out
