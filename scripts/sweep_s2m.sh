#!/bin/bash
out=gpurun_out/sweep_s2m.txt; : > $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
for occ in 6 5 4; do
  echo "== sr_lo occ $occ" >> $out
  VLOAM_LO_ASSOC_OCC=$occ timeout 300 python bench.py --workload sr_lo --legs device --steps 20 2>>gpurun_out/sweep_s2m.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f lo_associate %.1f'%(d['value'],d['ms_per_step'],d['kernels']['lo_associate']['avg_us']))" >> $out 2>&1
done
VLOAM_LO_ASSOC_OCC=4 python -m pytest tests/test_gpu_lidar.py -x -q 2>&1 | tail -2 >> $out
for occ in 6 4; do
  echo "== sr_lo_lm occ $occ" >> $out
  VLOAM_LO_ASSOC_OCC=$occ timeout 400 python bench.py --workload sr_lo_lm --legs device --steps 20 2>>gpurun_out/sweep_s2m.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']))
print('  '+' '.join('%s=%.0f'%(k,v['avg_us']) for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['avg_us']*kv[1]['launches'])[:12]))
" >> $out 2>&1
done
cat $out
