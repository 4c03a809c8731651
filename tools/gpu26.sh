set -x
mkdir -p gpurun_out
NBC_TIMING=1 timeout 900 python bench.py --workload cli --steps 2 --warmup 1 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; echo "bench cli exit $?"
tail -n 3 gpurun_out/bench_cli.err; grep -v "^{" gpurun_out/bench_cli.json; grep "^{" gpurun_out/bench_cli.json | cut -c1-200
NBC_TIMING=1 timeout 900 python bench.py --workload cli --steps 1 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err; echo "bench cli exit $?"
grep -v "^{" gpurun_out/bench_cli256.json; grep "^{" gpurun_out/bench_cli256.json | cut -c1-200
