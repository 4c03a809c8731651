# Round 2, GPU call 12: 3x3 halo kernel for layer1 (A/B NBC_HALO3), tests.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv" > gpurun_out/t_conv.log 2>&1; echo "pytest conv exit $?"
tail -n 15 gpurun_out/t_conv.log
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 8 gpurun_out/t_all.log
for h in 1 0; do
  NBC_HALO3=$h timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_halo3_$h.txt 2>&1
  echo "halo3=$h"; grep -E "layer1.*conv2|TOTAL" gpurun_out/layers_halo3_$h.txt
done
