# programmatic dependent launch between the conv kernels, and the chunk size of the engine
set -x
mkdir -p gpurun_out
NBC_PDL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k 'conv_tc_many_tiles or conv_dual or conv_full_width or ragged or engine' > gpurun_out/t_pdl.log 2>&1
echo "pytest exit $?"; grep -E 'passed|failed|Error' gpurun_out/t_pdl.log | head -5
for pdl in 0 1; do for n in 8 16; do
  NBC_PDL=$pdl timeout 200 python tools/prof_forward.py $n 624 1024 20 2>&1 | tail -n 1
done; done
for pdl in 0 1; do for c in 8 16; do
  echo "bench pdl=$pdl chunk=$c"
  NBC_PDL=$pdl NBC_CHUNK=$c timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl${pdl}_chunk$c.json 2> gpurun_out/bench_pdl${pdl}_chunk$c.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_pdl${pdl}_chunk$c.json') if l.startswith('{')][-1])
print('value %.1f e2e %.1f roofline %.3f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))
PY
done; done
