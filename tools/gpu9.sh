set -x
mkdir -p gpurun_out
PROF="python tools/prof_one_conv.py 256 1024 1 1 0 8 128 128 3"
timeout 120 $PROF > gpurun_out/plain9.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o /tmp/prof_conv3 $PROF > gpurun_out/ncu9.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/prof_conv3.ncu-rep --page source --csv > gpurun_out/prof_conv3nores_source.csv 2>/dev/null
ncu -i /tmp/prof_conv3.ncu-rep --page raw --csv > gpurun_out/prof_conv3nores_raw.csv 2>/dev/null
ls -la gpurun_out/
