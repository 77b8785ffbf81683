#!/bin/bash
# ncu artefacts for profiles/: launch list (every launch of 3 steady-state scans) and one --set full capture of one scan,
# for the default workload (configs[2]) at 64 streams per launch, one handle.  usage: scripts/gpu_profile.sh TAG
tag=${1:-r01c}
CMD="python bench.py --workload sr_lo_lm --legs device --batch 64 --handles 1 --steps 4 --warmup 3"
# one steady-state scan = 34 launches (10 SR + 6 LO + 18 LM); skip the 3 warm-up scans
ncu --metrics gpu__time_duration.sum --clock-control none -s 105 -c 102 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_launches.log
ncu --set full --clock-control none --import-source on -s 139 -c 34 -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
ls -la gpurun_out/${tag}*
