set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "train or predict_pipeline" > gpurun_out/t_train.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/t_train.log
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
cut -c1-260 gpurun_out/bench_train.json
