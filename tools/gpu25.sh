set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "pipeline or engine" > gpurun_out/t_pipe.log 2>&1; echo "pytest exit $?"
grep -v "Warning\|warn" gpurun_out/t_pipe.log | tail -n 30
nproc; free -g | head -2; df -h /dev/shm | tail -1
timeout 900 python bench.py --workload cli --steps 2 --warmup 1 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; echo "bench cli exit $?"
tail -n 5 gpurun_out/bench_cli.err; cat gpurun_out/bench_cli.json
