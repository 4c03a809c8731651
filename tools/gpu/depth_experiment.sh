set -x
mkdir -p gpurun_out
for cfg in "8 16" "16 16" "16 32"; do
  set -- $cfg
  echo "bench pdl=1 chunk=$1 depth=$2"
  NBC_PDL=1 NBC_CHUNK=$1 NBC_STAGE_DEPTH=$2 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c$1_d$2.json 2> gpurun_out/bench_c$1_d$2.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_c$1_d$2.json') if l.startswith('{')][-1])
print('value %.1f e2e %.1f roofline %.3f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))
PY
done
