set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "pipeline" > gpurun_out/t_pipe.log 2>&1; echo "pytest exit $?"
grep -v "Warning\|warn" gpurun_out/t_pipe.log | tail -n 5
for rep in 1 2; do
NBC_DEBUG_HANG=80 NBC_TIMING=1 timeout 150 python bench.py --workload cli --steps 3 --warmup 1 --batch 256 > gpurun_out/bench_cli256_$rep.json 2> gpurun_out/bench_cli256_$rep.err; echo "bench cli exit $?"
grep -v "^{" gpurun_out/bench_cli256_$rep.json | tail -n 4; grep "^{" gpurun_out/bench_cli256_$rep.json | cut -c1-200
tail -n 12 gpurun_out/bench_cli256_$rep.err | grep -v "^$"
done
NBC_DEBUG_HANG=80 NBC_TIMING=1 timeout 150 python bench.py --workload cli --steps 3 --warmup 1 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; echo "bench cli exit $?"
grep -v "^{" gpurun_out/bench_cli.json | tail -n 3; grep "^{" gpurun_out/bench_cli.json | cut -c1-200
