# Round 2, GPU call 22: predict.py CLI wall clock at engine chunk 8 / 16 (same box), with the pipeline's own timing.
set -x
mkdir -p gpurun_out
for c in 16 8 16 8; do
  NBC_TIMING=1 NBC_CHUNK=$c NBC_DEBUG_HANG=200 timeout 400 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli_chunk$c.json 2> gpurun_out/bench_cli_chunk$c.err
  echo "chunk $c"; grep "folder pipeline timing" gpurun_out/bench_cli_chunk$c.json | tail -2 | cut -c1-300
  grep "^{" gpurun_out/bench_cli_chunk$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f img/s, %.0f ms per step' % (d['value'], d['ms_per_step']))"
done
