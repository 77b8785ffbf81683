#!/bin/bash
# r02z (end of round 2): the last GPU calls, as run — full GPU suite, smoke(), front-end timing (synchronous and pipelined legs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -6
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python scripts/gpu_frontend_timing.py 32 2>&1 | tail -1      # -> profiles/r02z_frontend_pipelined_b32.json
timeout 60 python scripts/gpu_frontend_timing.py 1 2>&1 | tail -1        # -> profiles/r02z_frontend_pipelined_b1.json
# configs[3] with the image front end in the step:   python bench.py --workload vloam --steps 20 --warmup 3   -> profiles/r02z_bench_vloam.json
# the line of record:                                python bench.py                                          -> profiles/r02z_bench_default.json
