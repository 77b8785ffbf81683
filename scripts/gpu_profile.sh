#!/bin/bash
# ncu artefacts for profiles/: launch list (every launch of 3 steady-state scans) and one --set full capture of one scan,
# for the default workload (configs[2]) at 64 streams per launch, one handle.
CMD="python bench.py --workload sr_lo_lm --legs device --batch 64 --handles 1 --steps 4 --warmup 3"
# one scan = 32 launches (10 SR + 6 LO + 16 LM); the map seeding launches nothing but lm_init_state; skip the 3 warm-up scans
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 96 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_launches.log 2>&1
tail -2 gpurun_out/r01b_launches.log
ncu --set full --clock-control none --import-source on -s 132 -c 32 -o gpurun_out/r01b_full $CMD > gpurun_out/r01b_full.log 2>&1
tail -2 gpurun_out/r01b_full.log
ls -la gpurun_out/r01b*
