# Round 2, GPU call 5: cursor-based tile walk (A/B compact on/off, chunk 8/16) + K5 fusion, on ONE box.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
for c in 1 0 1 0; do
  NBC_RAGGED_COMPACT=$c timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown_compact$c.txt 2>&1
  echo "compact=$c"; grep -E "network|K5|whole" gpurun_out/step_breakdown_compact$c.txt
done
for c in 1 0; do
  NBC_CHUNK=16 NBC_RAGGED_COMPACT=$c timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown_chunk16_compact$c.txt 2>&1
  echo "chunk16 compact=$c"; grep -E "network|whole|K5" gpurun_out/step_breakdown_chunk16_compact$c.txt
done
for cfg in "8 1" "8 0" "16 1" "16 0"; do
  set -- $cfg
  NBC_CHUNK=$1 NBC_RAGGED_COMPACT=$2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c$1_k$2.json 2> gpurun_out/bench_c$1_k$2.err
  python - $1 $2 <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/bench_c%s_k%s.json' % (sys.argv[1], sys.argv[2])) if l.startswith('{')][-1])
print('chunk', sys.argv[1], 'compact', sys.argv[2], 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'])
PY
done
