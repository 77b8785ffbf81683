#!/bin/bash
# r02y: staged min-eigen tiles, one-atomic-per-CTA candidate append: parity, timing, sanitizers, ncu
mkdir -p gpurun_out
tag=${1:-r02y}
timeout 600 python -m pytest tests/test_vo_frontend.py tests/test_gpu_host_api.py tests/test_adapter_stubs.py tests/test_gpu_vo.py -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
timeout 200 python scripts/gpu_frontend_timing.py 32 > gpurun_out/${tag}_frontend.json 2> gpurun_out/${tag}_frontend.err; cat gpurun_out/${tag}_frontend.json; tail -3 gpurun_out/${tag}_frontend.err
timeout 200 python scripts/gpu_frontend_timing.py 1 > gpurun_out/${tag}_frontend_b1.json 2> gpurun_out/${tag}_frontend_b1.err; cat gpurun_out/${tag}_frontend_b1.json
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_vo_frontend.py -m gpu -q -x -k "detector or chain" > gpurun_out/${tag}_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -4
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_vo_frontend.py -m gpu -q -x -k "detector or chain" > gpurun_out/${tag}_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"vo_orb_describe|vo_bf_match|vo_min_eigen|vo_corner|vo_select" -s 14 -c 5 -o gpurun_out/${tag}_frontend_full python scripts/gpu_frontend_timing.py 32 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
