set -x
mkdir -p gpurun_out
for o in 4 2 1 0; do
echo "NBC_WGRAD_OVER=$o"
NBC_WGRAD_OVER=$o timeout 300 python bench.py --workload train --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | cut -c1-140
done
