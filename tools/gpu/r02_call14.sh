# Round 2, GPU call 14: 3x3 halo kernel (final form) + teacher-forced training forward test: full suite, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 8 gpurun_out/t_all.log
timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown.txt 2>&1; tail -n 7 gpurun_out/step_breakdown.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])
PY
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"; cut -c1-160 gpurun_out/bench_train.json
