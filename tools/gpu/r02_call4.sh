# Round 2, GPU call 4: A/B of the compact ragged tile enumeration and of the chunk size, on ONE box.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_base.txt 2>&1; tail -n 1 gpurun_out/layers_base.txt
for c in 1 0 1 0; do
  NBC_RAGGED_COMPACT=$c timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown_compact$c.txt 2>&1
  echo "compact=$c"; grep -E "network|whole" gpurun_out/step_breakdown_compact$c.txt
done
for c in 1 0; do
  NBC_CHUNK=16 NBC_RAGGED_COMPACT=$c timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown_chunk16_compact$c.txt 2>&1
  echo "chunk16 compact=$c"; grep -E "network|whole|K5" gpurun_out/step_breakdown_chunk16_compact$c.txt
done
for c in 1 0; do
  NBC_RAGGED_COMPACT=$c timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_compact$c.json 2> gpurun_out/bench_compact$c.err
  python - $c <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/bench_compact%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print('compact', sys.argv[1], 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'])
PY
done
NBC_CHUNK=16 NBC_RAGGED_COMPACT=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_chunk16.json 2> gpurun_out/bench_chunk16.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_chunk16.json') if l.startswith('{')][-1])
print('chunk16 compact0 value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'])
PY
