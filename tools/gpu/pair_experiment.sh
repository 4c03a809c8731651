# CTA-pair (cta_group::2) convolution: correctness first, then the per-layer table for every layer class.
set -x
mkdir -p gpurun_out
NBC_CTA2=15 timeout 120 python tools/gpu/pair_debug.py > gpurun_out/pair_debug.log 2>&1; rc=$?
echo "pair_debug exit $rc"; cat gpurun_out/pair_debug.log | tail -n 40
if [ $rc -ne 0 ]; then exit 0; fi
NBC_CTA2=15 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k 'conv_all_shapes and 1-shape or conv_tc_many_tiles or conv_dual or conv_full_width or ragged' > gpurun_out/t_pair.log 2>&1
echo "pytest exit $?"; grep -E 'passed|failed|Error' gpurun_out/t_pair.log | head -5
for m in 0 1 2 4 8 15; do
  NBC_CTA2=$m timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_pair$m.txt 2>&1
  echo "mode $m"; tail -n 1 gpurun_out/layers_pair$m.txt
done
