set -x
mkdir -p gpurun_out
PROF="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $PROF > gpurun_out/plain20.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF > gpurun_out/ncu20.log 2>&1; echo "ncu exit $?"
tail -n 1 gpurun_out/plain20.log
