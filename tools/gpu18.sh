set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "pytest train+model exit $?"
tail -n 8 gpurun_out/t_train.log
timeout 600 python bench.py --workload train --steps 3 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
tail -n 3 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
