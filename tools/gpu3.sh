set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x --no-header -p no:cacheprovider -s -k "stem or model or batch_equals" > gpurun_out/t_stem_model.log 2>&1; echo "pytest exit $?"
timeout 300 python tools/layer_profile.py 1 624 1024 > gpurun_out/layers_n1_624.txt 2>&1; echo "lp1 exit $?"
timeout 300 python tools/layer_profile.py 8 1024 1024 > gpurun_out/layers_n8_1024.txt 2>&1; echo "lp8 exit $?"
timeout 300 python tools/layer_profile.py 2 1024 1024 2 > gpurun_out/layers_n2_1024_mma.txt 2>&1; echo "lp2mma exit $?"
PROF="python tools/layer_profile.py 4 1024 1024"
timeout 300 $PROF > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 273 -c 40 -o /tmp/prof_conv_tc $PROF > gpurun_out/ncu3.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/prof_conv_tc.ncu-rep --page raw --csv > gpurun_out/prof_conv_tc_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv_tc.ncu-rep --page details --csv > gpurun_out/prof_conv_tc_details.csv 2>/dev/null
ls -la /tmp/prof_conv_tc.ncu-rep gpurun_out/
