set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -s -x > gpurun_out/t_train.log 2>&1; echo "pytest exit $?"
tail -n 120 gpurun_out/t_train.log
