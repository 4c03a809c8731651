# Round 2, GPU call 7: K1 per chunk (batched), K3 with the shared x pass, K5 fusion -- tests, breakdown, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 3 gpurun_out/t_all.log
timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown.txt 2>&1; tail -n 7 gpurun_out/step_breakdown.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value %.1f' % d['value'], 'e2e', d['e2e'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step'])
PY
timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_base.txt 2>&1; tail -n 1 gpurun_out/layers_base.txt
timeout 600 python bench.py --workload batch32 --no-cpu-baseline > gpurun_out/bench_batch32.json 2> gpurun_out/bench_batch32.err; echo "bench batch32 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_batch32.json') if l.startswith('{')][-1])
print('batch32 value %.1f' % d['value'], 'e2e', d['e2e'], 'clocks', d['clocks'])
PY
NBC_DEBUG_HANG=200 timeout 400 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err; echo "bench cli exit $?"
grep "^{" gpurun_out/bench_cli256.json | cut -c1-300
NBC_COMBINED=0 NBC_DEBUG_HANG=200 timeout 400 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256_nocombined.json 2> /dev/null
grep "^{" gpurun_out/bench_cli256_nocombined.json | cut -c1-200
