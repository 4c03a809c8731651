# Round 2, GPU call 9: run-based K5 (bit-exact tests, A/B against the per-pixel kernels), bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 12 gpurun_out/t_all.log
NBC_CCL_RUNS=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "ccl" > gpurun_out/t_ccl_pixel.log 2>&1; echo "pytest ccl (pixel kernels) exit $?"
tail -n 2 gpurun_out/t_ccl_pixel.log
for r in 1 0 1 0; do
  NBC_CCL_RUNS=$r timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown_runs$r.txt 2>&1
  echo "ccl runs=$r"; grep -E "network|K5|whole" gpurun_out/step_breakdown_runs$r.txt
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
