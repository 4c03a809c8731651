# Round 2, GPU call 20: chunk size with the per-chunk K1 (value and e2e), same box.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "ccl_random" 2>&1 | tail -2
for c in 8 16 12 8 16; do
  NBC_CHUNK=$c timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err
  python - $c <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/bench_chunk%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print('chunk', sys.argv[1], 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks']['sm_mhz'])
PY
done
