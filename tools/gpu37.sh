set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x -k "dual or golden or full_size or ragged or engine or trained_like or pipeline" -s > gpurun_out/t_dual.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|Error|error|golden small|1024x1024 |611x1024 |trained-like" gpurun_out/t_dual.log | tail -n 30
timeout 300 python tools/layer_profile.py 8 1024 1024 > gpurun_out/layers_n8_1024.txt 2>&1; grep -E "downsample|TOTAL" gpurun_out/layers_n8_1024.txt
timeout 300 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_n8_624.txt 2>&1; tail -n 1 gpurun_out/layers_n8_624.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 2 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench_nocpu.json
