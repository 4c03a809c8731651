set -x
mkdir -p gpurun_out
for rep in 1 2; do
NBC_DEBUG_HANG=60 NBC_TIMING=1 timeout 150 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256_$rep.json 2> gpurun_out/bench_cli256_$rep.err; echo "bench cli exit $?"
grep -v "^{" gpurun_out/bench_cli256_$rep.json | tail -n 8; grep "^{" gpurun_out/bench_cli256_$rep.json | cut -c1-200
tail -n 120 gpurun_out/bench_cli256_$rep.err | grep -v "^$" | head -150
done
