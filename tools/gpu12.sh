set -x
mkdir -p gpurun_out
PROF="python bench.py --workload train --batch 4 --steps 1 --warmup 1"
timeout 300 $PROF > gpurun_out/plain12.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 640 --csv --log-file gpurun_out/launches_train.csv $PROF > gpurun_out/ncu12.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/plain12.log
