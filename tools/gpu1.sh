set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug exit $?"
NBC_TEST_IMPLS=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/t_kernels_mma.log 2>&1; echo "kernels(mma) exit $?"
NBC_TEST_IMPLS=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "conv" > gpurun_out/t_kernels_tc.log 2>&1; echo "kernels(tc) exit $?"
NBC_TEST_IMPLS=2 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --no-header -p no:cacheprovider -k "golden_small or batch_equals" > gpurun_out/t_model_mma.log 2>&1; echo "model(mma) exit $?"
tail -5 gpurun_out/tc_debug.log gpurun_out/t_kernels_mma.log gpurun_out/t_kernels_tc.log gpurun_out/t_model_mma.log
