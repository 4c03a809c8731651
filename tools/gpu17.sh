set -x
mkdir -p gpurun_out
PROF="python bench.py --workload train --batch 8 --steps 1 --warmup 1"
timeout 300 $PROF > gpurun_out/plain17.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 912 -c 912 --csv --log-file gpurun_out/launches_train.csv $PROF > gpurun_out/ncu17.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/plain17.log
