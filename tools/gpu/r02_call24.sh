# Round 2, GPU call 24: folder pipeline with the mapping reader -- pipeline tests + predict.py CLI wall clock.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "pipeline or engine" 2>&1 | tail -3
for i in 1 2; do
  NBC_TIMING=1 NBC_DEBUG_HANG=200 timeout 400 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err
  grep "folder pipeline timing" gpurun_out/bench_cli256.json | tail -1 | cut -c1-200
  grep "^{" gpurun_out/bench_cli256.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f img/s, %.0f ms per step' % (d['value'], d['ms_per_step']))"
done
