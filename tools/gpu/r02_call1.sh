# Round 2, GPU call 1: never-run kernels, the fp16-default parity gates, bench lines, two prepared experiments.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
# (1) general-ratio resize (ungated now) + the whole suite
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 15 gpurun_out/t_all.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -s -k 'north_star or golden_small or full_size or trained_like or saturates or pipeline_matches' > gpurun_out/t_parity.log 2>&1; echo "parity exit $?"
grep -h "^\[" gpurun_out/t_parity.log | head -60
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
# (2) stem halo kernel
NBC_TEST_EXPERIMENTAL=1 NBC_STEM_HALO=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k 'stem or ragged or engine or north_star' > gpurun_out/t_stem_halo.log 2>&1
echo "stem halo exit $?"; tail -n 12 gpurun_out/t_stem_halo.log
# (3) bench lines
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?"; cut -c1-300 gpurun_out/bench_bf16.json
timeout 600 python bench.py --workload batch32 --no-cpu-baseline > gpurun_out/bench_batch32.json 2> gpurun_out/bench_batch32.err; echo "bench batch32 exit $?"; cut -c1-600 gpurun_out/bench_batch32.json
# (4) per-layer tables: baseline, residual expansions on 128-column tiles, stem halo
timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_base.txt 2>&1; tail -n 1 gpurun_out/layers_base.txt
NBC_RES_BN=128 timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_resbn128.txt 2>&1; tail -n 1 gpurun_out/layers_resbn128.txt
NBC_STEM_HALO=1 timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_stem_halo.txt 2>&1; head -n 3 gpurun_out/layers_stem_halo.txt; tail -n 1 gpurun_out/layers_stem_halo.txt
timeout 200 python tools/layer_profile.py 32 1024 1024 > gpurun_out/layers_n32_1024.txt 2>&1; tail -n 1 gpurun_out/layers_n32_1024.txt
# (5) training bench on the final code
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
cut -c1-400 gpurun_out/bench_train.json
