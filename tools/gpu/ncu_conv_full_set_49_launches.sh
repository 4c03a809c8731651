set -x
mkdir -p gpurun_out
PROF="python tools/prof_forward.py 8 624 1024 3"
timeout 300 $PROF > gpurun_out/plain38.log 2>&1; echo "plain exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 98 -c 49 -o /tmp/conv_full $PROF > gpurun_out/ncu38.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
ls -la /tmp/conv_full.ncu-rep gpurun_out/conv_full_raw.csv
