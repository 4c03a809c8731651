set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/t_train.log
timeout 900 python bench.py --workload train --steps 3 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
cat gpurun_out/bench_train.json; tail -n 5 gpurun_out/bench_train.err
nvidia-smi --query-gpu=memory.used --format=csv
