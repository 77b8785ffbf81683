set -x
mkdir -p gpurun_out/d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d/bench_default.json 2> gpurun_out/d/bench_default.err
timeout 900 python bench.py --workload vloam --steps 6 --warmup 3 --cpu-scans 60 > gpurun_out/d/bench_vloam.json 2> gpurun_out/d/bench_vloam.err
timeout 900 python bench.py --workload sr_lo_lm --steps 6 --warmup 3 --cpu-scans 60 > gpurun_out/d/bench_map.json 2> gpurun_out/d/bench_map.err
ls -la gpurun_out/d
