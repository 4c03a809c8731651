set -x
mkdir -p gpurun_out
PROF="python tools/prof_forward.py 8 624 1024 3"
timeout 300 $PROF > gpurun_out/plain22.log 2>&1; echo "plain exit $?"
# (1) DRAM traffic + duration of the 53 conv launches of the LAST forward (skip 2 x 53 warm launches)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:conv_tc_ -s 106 -c 53 --csv --log-file gpurun_out/conv_traffic.csv $PROF > gpurun_out/ncu22a.log 2>&1; echo "ncu a exit $?"
# (2) full-set capture of two representative launches: the head 3x3 conv (compute-bound) and a layer3 conv3 (HBM-bound)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 106 -c 53 -o /tmp/conv_full $PROF > gpurun_out/ncu22b.log 2>&1; echo "ncu b exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
ncu -i /tmp/conv_full.ncu-rep --page details --csv > gpurun_out/conv_full_details.csv 2>/dev/null
ls -la /tmp/conv_full.ncu-rep gpurun_out/
