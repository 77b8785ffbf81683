set -x
mkdir -p gpurun_out/f
timeout 600 python -m pytest tests/test_gpu_shard.py tests/test_gpu_vo.py -m gpu -x -q > gpurun_out/f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 3 --batch 8 --handles 1 --parallelism point --cpu-scans 20 > gpurun_out/f/point_b8_n2.json 2> gpurun_out/f/point_b8_n2.err
timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 30 --warmup 3 --batch 64 --handles 1 --parallelism point --cpu-scans 20 > gpurun_out/f/point_b64_n2.json 2> gpurun_out/f/point_b64_n2.err
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 30 --warmup 3 --cpu-scans 20 > gpurun_out/f/stream_n2.json 2> gpurun_out/f/stream_n2.err
ls -la gpurun_out/f
