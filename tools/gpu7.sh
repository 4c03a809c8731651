set -x
mkdir -p gpurun_out
for cfg in "256 1024 1 1 1 8 128 128" "256 1024 1 1 0 8 128 128" "1024 256 1 1 0 8 128 128" "64 256 1 1 1 8 256 256" "512 2048 1 1 1 8 128 128" "256 256 3 2 0 8 128 128" "2048 512 3 1 0 8 128 128"; do
  timeout 120 python tools/prof_one_conv.py $cfg >> gpurun_out/one_conv.txt 2>&1
done
cat gpurun_out/one_conv.txt
PROF="python tools/prof_one_conv.py 256 1024 1 1 1 8 128 128 3"
timeout 120 $PROF > gpurun_out/plain7.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o /tmp/prof_conv3 $PROF > gpurun_out/ncu7.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/prof_conv3.ncu-rep --page details --csv > gpurun_out/prof_conv3_details.csv 2>/dev/null
ncu -i /tmp/prof_conv3.ncu-rep --page source --csv > gpurun_out/prof_conv3_source.csv 2>/dev/null
ncu -i /tmp/prof_conv3.ncu-rep --page raw --csv > gpurun_out/prof_conv3_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/prof_conv3.ncu-rep
