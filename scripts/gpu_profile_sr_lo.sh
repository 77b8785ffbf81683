set -x
mkdir -p gpurun_out/g
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/g/launches.csv python bench.py --legs device --steps 4 --warmup 2 --batch 128 --handles 1 > gpurun_out/g/launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"^sr_|^lo_" --launch-skip 12 -c 16 -o gpurun_out/g/full_sr_lo python bench.py --legs device --steps 1 --warmup 1 --batch 128 --handles 1 > gpurun_out/g/ncu_full.log 2>&1
ncu -i gpurun_out/g/full_sr_lo.ncu-rep --page raw --csv > gpurun_out/g/full_sr_lo_raw.csv 2>/dev/null
du -sh gpurun_out/g/*
