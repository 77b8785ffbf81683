set -x
mkdir -p gpurun_out/a
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a/pytest.log
for cfg in "1 128 4" "2 128 4" "4 128 4" "1 128 5" "1 128 6" "1 256 4" "2 256 4" "4 256 4"; do
  set -- $cfg
  VLOAM_LO_ASSOC_OCC=$3 timeout 300 python bench.py --legs device --steps 10 --warmup 3 --handles $1 --batch $2 > gpurun_out/a/dev_h$1_b$2_o$3.json 2> gpurun_out/a/dev_h$1_b$2_o$3.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sr_less_flat_voxel|lo_associate|sr_pick_features|lo_solve" -c 8 -o gpurun_out/a/full_hot python bench.py --legs device --steps 1 --warmup 1 --batch 128 > gpurun_out/a/ncu_full.log 2>&1
ls -la gpurun_out/a
