set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "wgrad" > gpurun_out/t_wgrad.log 2>&1; echo "pytest wgrad exit $?"
tail -n 25 gpurun_out/t_wgrad.log
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "pytest train exit $?"
tail -n 25 gpurun_out/t_train.log
timeout 600 python bench.py --workload train --steps 3 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
tail -n 3 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
