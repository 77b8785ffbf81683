set -x
mkdir -p gpurun_out/e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e/pytest.log
timeout 300 python bench.py --steps 30 --warmup 3 --batch 8 --handles 1 --cpu-scans 20 > gpurun_out/e/b8_n1.json 2> gpurun_out/e/b8_n1.err
ls -la gpurun_out/e
