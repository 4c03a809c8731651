# First GPU call of the next round: everything round 1 built after its GPU budget ran out, then the standard verification.
#   gpurun --timeout 1500 -- 'bash tools/gpu/round2_first_call.sh'
set -x
mkdir -p gpurun_out
# (1) the general-ratio resize kernel (never run on a GPU so far)
NBC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k general_ratio > gpurun_out/t_general.log 2>&1
echo "general-ratio exit $?"; tail -n 5 gpurun_out/t_general.log
# (2) full suite + smoke
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
# (3) bench lines: predict in both storage precisions, the CLI with the native PNG encoder, training
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json
timeout 600 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench fp16 exit $?"; cut -c1-400 gpurun_out/bench_fp16.json
NBC_DEBUG_HANG=100 timeout 300 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err; echo "bench cli exit $?"
grep "^{" gpurun_out/bench_cli256.json | cut -c1-260
NBC_COMBINED=0 NBC_DEBUG_HANG=100 timeout 300 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256_nocombined.json 2> /dev/null
grep "^{" gpurun_out/bench_cli256_nocombined.json | cut -c1-260
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
cut -c1-300 gpurun_out/bench_train.json
# (4) the ncu --set full capture of all 49 conv launches that round 1's budget cut off (allow 10 minutes)
PROF="python tools/prof_forward.py 8 624 1024 3"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 98 -c 49 -o /tmp/conv_full $PROF > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
python tools/ncu_summarise.py gpurun_out/conv_full_raw.csv gpurun_out/ncu_conv_tc_full.txt gpurun_out/ncu_traffic.json
# (5) residual 1x1 expansions on 128-column tiles (5 operand stages instead of 3)
for v in 0 128; do NBC_RES_BN=$v timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_resbn$v.txt 2>&1; echo "NBC_RES_BN=$v"; tail -n 1 gpurun_out/layers_resbn$v.txt; done
# (6) the stem halo kernel (overlapping no-swizzle windows; never run on a GPU so far): its own test, the model-level tests
# and the per-layer table with it switched on
NBC_TEST_EXPERIMENTAL=1 NBC_STEM_HALO=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k 'stem or ragged or engine or model_parity' > gpurun_out/t_stem_halo.log 2>&1
echo "stem halo exit $?"; tail -n 5 gpurun_out/t_stem_halo.log
NBC_STEM_HALO=1 timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_stem_halo.txt 2>&1; head -n 3 gpurun_out/layers_stem_halo.txt; tail -n 1 gpurun_out/layers_stem_halo.txt
