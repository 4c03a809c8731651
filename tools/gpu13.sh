set -x
mkdir -p gpurun_out; rm -f gpurun_out/one_conv.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv or ragged or golden_small or stem" > gpurun_out/t_conv.log 2>&1; echo "pytest exit $?"
tail -n 15 gpurun_out/t_conv.log
for cfg in "256 1024 1 1 1 8 128 128" "256 1024 1 1 0 8 128 128" "1024 256 1 1 0 8 128 128" "64 256 1 1 1 8 256 256" "512 2048 1 1 1 8 128 128" "256 256 3 2 0 8 128 128" "2048 512 3 1 0 8 128 128"; do
  timeout 120 python tools/prof_one_conv.py $cfg >> gpurun_out/one_conv.txt 2>&1
done
cat gpurun_out/one_conv.txt
timeout 300 python tools/layer_profile.py 8 1024 1024 > gpurun_out/layers_n8_1024.txt 2>&1; tail -n 1 gpurun_out/layers_n8_1024.txt
timeout 300 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_n8_624.txt 2>&1; tail -n 1 gpurun_out/layers_n8_624.txt
