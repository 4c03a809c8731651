set -x
mkdir -p gpurun_out
# (1) HBM-side kernels of the predict path: K1, K3, K5 with the full metric set (one chunk of 8)
PROF="python bench.py --batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $PROF > gpurun_out/plain41.log 2>&1; echo "plain exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"resize4x|upsample_argmax|ccl_|maxpool|head1x1|stem_pad|trim_rows" -s 30 -c 30 -o /tmp/small_full $PROF > gpurun_out/ncu41a.log 2>&1; echo "ncu a exit $?"
ncu -i /tmp/small_full.ncu-rep --page raw --csv > gpurun_out/small_full_raw.csv 2>/dev/null
# (2) launch list of the predict bench (final code)
PROF2="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF2 > gpurun_out/ncu41b.log 2>&1; echo "ncu b exit $?"
# (3) training step: launch list of one step + full set of the weight-gradient kernel
PROF3="python bench.py --workload train --batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $PROF3 > gpurun_out/plain41c.log 2>&1; echo "plain train exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv $PROF3 > gpurun_out/ncu41c.log 2>&1; echo "ncu c exit $?"
timeout 900 ncu --set full --clock-control none -k regex:wgrad_tc_kernel -s 60 -c 12 -o /tmp/wgrad_full $PROF3 > gpurun_out/ncu41d.log 2>&1; echo "ncu d exit $?"
ncu -i /tmp/wgrad_full.ncu-rep --page raw --csv > gpurun_out/wgrad_full_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv
