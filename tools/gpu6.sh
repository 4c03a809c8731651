set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/t_all.log 2>&1; echo "pytest exit $?"
timeout 300 python tools/layer_profile.py 8 1024 1024 > gpurun_out/layers_n8_1024.txt 2>&1; echo "lp8 exit $?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; grep -E "passed|failed" gpurun_out/t_all.log; tail -n 2 gpurun_out/layers_n8_1024.txt
