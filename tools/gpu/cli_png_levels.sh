set -x
mkdir -p gpurun_out
for cfg in "NBC_PNG_LEVEL=0 NBC_COMBINED=0" "NBC_PNG_LEVEL=0 NBC_COMBINED=1" "NBC_PNG_LEVEL=1 NBC_COMBINED=0"; do
echo "$cfg"
env $cfg NBC_DEBUG_HANG=150 NBC_TIMING=1 timeout 200 python bench.py --workload cli --batch 1024 --steps 1 --warmup 1 2>/dev/null | grep -v "folders" | cut -c1-230 | tail -n 2
done
