# ncu evidence for the final code: (1) --set full over the 49 conv launches of one network pass, (2) launch list of the bench command
set -x
mkdir -p gpurun_out
PROF="python tools/prof_forward.py 8 624 1024 3"
timeout 100 $PROF > gpurun_out/plain_fwd.log 2>&1; echo "plain exit $?"; tail -n 1 gpurun_out/plain_fwd.log
timeout 230 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 98 -c 49 -o /tmp/conv_full $PROF > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
ls -la /tmp/conv_full.ncu-rep gpurun_out/conv_full_raw.csv
PROF2="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
wc -l gpurun_out/launches_predict.csv
