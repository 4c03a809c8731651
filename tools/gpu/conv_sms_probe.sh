# does leaving SMs free for the side streams (K1, K3, K5) pay?  value with the conv kernels on NBC_CONV_SMS SMs
mkdir -p gpurun_out
NBC_CONV_SMS=${1:-132} timeout 55 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_sms.json 2> gpurun_out/bench_sms.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_sms.json') if l.startswith('{')][-1])
print('value %.1f roofline %.3f clocks %s' % (d['value'], d['roofline']['frac'], d['clocks']))
PY
