#!/bin/bash
out=gpurun_out/sweep_s2e.txt; : > $out
python -m pytest tests/test_gpu_mapping.py -x -q 2>&1 | tail -2 >> $out
for cs in 1 2 4 8 1 8; do
  echo "== cluster $cs" >> $out
  VLOAM_LM_CLUSTER=$cs timeout 300 python bench.py --workload sr_lo_lm --legs device --steps 40 2>>gpurun_out/sweep_s2e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f lm_solve %.1f'%(d['value'],d['ms_per_step'],d['kernels']['lm_solve']['avg_us']))" >> $out 2>&1
done
cat $out
