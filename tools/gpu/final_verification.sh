set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
cut -c1-300 gpurun_out/bench_train.json
NBC_DEBUG_HANG=100 timeout 200 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err; echo "bench cli exit $?"
grep "^{" gpurun_out/bench_cli256.json | cut -c1-260
timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_final.txt 2>&1; tail -n 1 gpurun_out/layers_final.txt
