#!/bin/bash
# ncu launch list + full capture of one steady-state scan (64 streams, one handle): artefacts for profiles/
mkdir -p gpurun_out
tag=${1:-r02h}
CMD="python bench.py --workload sr_lo_lm --legs device --batch 64 --handles 1 --steps 3 --warmup 4"
# per scan: 10 SR + 6 LO + 18 LM launches = 34; skip the handle set-up launches and 4 warm-up scans, list 3 scans
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sr_|lo_|lm_|gn_" -s 140 -c 102 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_launches.log
ncu --set full --clock-control none --import-source on -k regex:"sr_|lo_|lm_|gn_" -s 174 -c 34 -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
ls -la gpurun_out/${tag}*
