set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "trained_like" -s > gpurun_out/t_parity2.log 2>&1; echo "pytest exit $?"
grep "trained-like\|passed\|failed" gpurun_out/t_parity2.log
