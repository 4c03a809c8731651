# Round 2, GPU call 11: ragged-aware head 1x1 / stem staging -- tests, breakdown, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown.txt 2>&1; tail -n 7 gpurun_out/step_breakdown.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
