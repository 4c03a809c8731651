set -x
mkdir -p gpurun_out
# BASELINE.json configs[4] at 1 GPU: predict.py over a 10 000-entry folder (hard links to a pool of unique 4096^2 BMPs)
NBC_DEBUG_HANG=500 NBC_TIMING=1 timeout 560 python bench.py --workload cli --batch 10000 --steps 1 --warmup 0 > gpurun_out/bench_cli10k.json 2> gpurun_out/bench_cli10k.err; echo "cli 10k exit $?"
grep -v "^{" gpurun_out/bench_cli10k.json | tail -n 4; grep "^{" gpurun_out/bench_cli10k.json | cut -c1-700
tail -n 3 gpurun_out/bench_cli10k.err
