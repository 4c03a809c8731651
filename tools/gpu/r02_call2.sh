# Round 2, GPU call 2: stem halo default + zero-band span verified, step breakdown, e2e A/B, fresh ncu evidence.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 4 gpurun_out/t_all.log
timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown.txt 2>&1; cat gpurun_out/step_breakdown.txt | tail -8
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json
NBC_ZERO_SPAN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nospan.json 2> gpurun_out/bench_nospan.err; echo "bench nospan exit $?"
python - <<'PY'
import json
for f in ('bench.json', 'bench_nospan.json'):
    d = json.loads([l for l in open('gpurun_out/' + f) if l.startswith('{')][-1])
    print(f, 'value %.1f' % d['value'], 'e2e', d['e2e'], 'clocks', d['clocks'])
PY
timeout 200 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_base.txt 2>&1; tail -n 1 gpurun_out/layers_base.txt
# ---- ncu (evidence only; nothing printed under ncu is a bench value) ------------------------------------------------
# (a) launch list of one bench step
PROF2="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF2 > gpurun_out/ncu_a.log 2>&1; echo "ncu a exit $?"
# (b) K1 / K3 / K5 / maxpool / head 1x1 with the full set (one chunk of 8)
PROF1="python bench.py --batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resize4x|upsample_argmax|ccl_|maxpool|head1x1|stem_pad|trim_rows" -s 30 -c 30 -o /tmp/small_full $PROF1 > gpurun_out/ncu_b.log 2>&1; echo "ncu b exit $?"
ncu -i /tmp/small_full.ncu-rep --page raw --csv > gpurun_out/small_full_raw.csv 2>/dev/null
# (c) every tensor-core conv launch of one network pass with the full set
PROF="python tools/prof_forward.py 8 624 1024 3"
timeout 100 $PROF > gpurun_out/plain_fwd.log 2>&1; echo "plain exit $?"; tail -n 1 gpurun_out/plain_fwd.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 100 -c 50 -o /tmp/conv_full $PROF > gpurun_out/ncu_c.log 2>&1; echo "ncu c exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
python tools/ncu_summarise.py gpurun_out/conv_full_raw.csv gpurun_out/ncu_conv_tc_full.txt gpurun_out/ncu_traffic.json
ls -la gpurun_out/*.csv
