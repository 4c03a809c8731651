set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider -s > gpurun_out/t_all.log 2>&1; echo "pytest exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
kill $SMI
timeout 600 python bench.py --workload batch32 --no-cpu-baseline > gpurun_out/bench_batch32.json 2> gpurun_out/bench_batch32.err; echo "bench32 exit $?"
PROF="python bench.py --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu1.log 2>&1; echo "ncu launches exit $?"
timeout 300 $PROF > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernelILi256 -s 40 -c 3 -o gpurun_out/prof_conv_tc $PROF > gpurun_out/ncu2.log 2>&1; echo "ncu full exit $?"
cat gpurun_out/bench.json; tail -n 3 gpurun_out/t_all.log
