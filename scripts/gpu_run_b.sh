set -x
mkdir -p gpurun_out/b
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b/pytest.log
for cfg in "1 128 6" "2 128 6" "1 128 8" "2 256 6"; do
  set -- $cfg
  VLOAM_LO_ASSOC_OCC=$3 timeout 300 python bench.py --legs device --steps 10 --warmup 3 --handles $1 --batch $2 > gpurun_out/b/dev_h$1_b$2_o$3.json 2> gpurun_out/b/dev_h$1_b$2_o$3.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lo_associate|lo_build_grid|sr_pack" -c 4 -o gpurun_out/b/full_lo python bench.py --legs device --steps 1 --warmup 0 --batch 128 > gpurun_out/b/ncu_full.log 2>&1
ls -la gpurun_out/b
