set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "augment" > gpurun_out/t_aug.log 2>&1; echo "pytest exit $?"
grep -v "Warning\|warn" gpurun_out/t_aug.log | tail -n 25
