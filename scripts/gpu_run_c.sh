set -x
mkdir -p gpurun_out/c
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c/pytest.log
for cfg in "1 128" "2 128" "2 256"; do
  set -- $cfg
  timeout 300 python bench.py --legs device --steps 10 --warmup 3 --handles $1 --batch $2 > gpurun_out/c/dev_h$1_b$2.json 2> gpurun_out/c/dev_h$1_b$2.err
done
for cfg in "1 8" "1 32" "2 64"; do
  set -- $cfg
  timeout 600 python bench.py --legs device --workload sr_lo_lm --steps 6 --warmup 3 --handles $1 --batch $2 > gpurun_out/c/map_h$1_b$2.json 2> gpurun_out/c/map_h$1_b$2.err
done
ls -la gpurun_out/c
