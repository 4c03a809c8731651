set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
cut -c1-400 gpurun_out/bench_ref.json
PROF="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF > gpurun_out/ncu36.log 2>&1; echo "ncu exit $?"
