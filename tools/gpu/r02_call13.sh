# Round 2, GPU call 13: 3x3 halo kernel, descriptor base-offset field off (0) / on (1).
set -x
mkdir -p gpurun_out
for bo in 0 1; do
  NBC_HALO3_BO=$bo timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "many_tiles or full_width" > gpurun_out/t_conv_bo$bo.log 2>&1; echo "bo=$bo pytest conv exit $?"
  grep -E "passed|failed|max err" gpurun_out/t_conv_bo$bo.log | head -5
done
