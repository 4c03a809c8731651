# Round 2, GPU call 10: verification of the final code + fresh evidence (bench lines, launch list, ncu --set full).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -rs > gpurun_out/t_all.log 2>&1; echo "pytest all exit $?"
tail -n 6 gpurun_out/t_all.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -s -k 'north_star or golden_small or full_size or trained_like or saturates or pipeline_matches' > gpurun_out/t_parity.log 2>&1; echo "parity exit $?"
grep -h "^\[north" gpurun_out/t_parity.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 2 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
timeout 600 python bench.py --workload batch32 > gpurun_out/bench_batch32.json 2> gpurun_out/bench_batch32.err; echo "bench batch32 exit $?"
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"
NBC_DEBUG_HANG=200 timeout 400 python bench.py --workload cli --steps 2 --warmup 1 --batch 256 > gpurun_out/bench_cli256.json 2> gpurun_out/bench_cli256.err; echo "bench cli exit $?"
timeout 200 python tools/layer_profile.py 16 624 1024 > gpurun_out/layers_base.txt 2>&1; tail -n 1 gpurun_out/layers_base.txt
timeout 200 python tools/layer_profile.py 32 1024 1024 > gpurun_out/layers_n32_1024.txt 2>&1; tail -n 1 gpurun_out/layers_n32_1024.txt
timeout 300 python tools/step_breakdown.py 64 > gpurun_out/step_breakdown.txt 2>&1; tail -n 7 gpurun_out/step_breakdown.txt
python - <<'PY'
import json
for f in ('bench', 'bench_reference', 'bench_batch32', 'bench_train', 'bench_cli256'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.2f' % d['value'], 'e2e', d.get('e2e') and round(d['e2e']['value'], 2), 'clocks', d.get('clocks'), 'roof', d.get('roofline') and (round(d['roofline']['achieved'], 1), round(d['roofline']['frac'], 3)), 'cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
# ---- ncu (evidence only) ----------------------------------------------------------------------------------------------
PROF2="python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_predict.csv $PROF2 > gpurun_out/ncu_a.log 2>&1; echo "ncu a exit $?"
PROF1="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resize4x|trim_rows|upsample_argmax|rows_|maxpool|head1x1|stem_pad" -s 14 -c 14 -o /tmp/small_full $PROF1 > gpurun_out/ncu_b.log 2>&1; echo "ncu b exit $?"
ncu -i /tmp/small_full.ncu-rep --page raw --csv > gpurun_out/small_full_raw.csv 2>/dev/null
PROF3="python bench.py --workload train --batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv $PROF3 > gpurun_out/ncu_d.log 2>&1; echo "ncu d exit $?"
timeout 900 ncu --set full --clock-control none -k regex:wgrad_tc_kernel -s 60 -c 12 -o /tmp/wgrad_full $PROF3 > gpurun_out/ncu_e.log 2>&1; echo "ncu e exit $?"
ncu -i /tmp/wgrad_full.ncu-rep --page raw --csv > gpurun_out/wgrad_full_raw.csv 2>/dev/null
PROF="python tools/prof_forward.py 16 624 1024 3"
timeout 100 $PROF > gpurun_out/plain_fwd.log 2>&1; echo "plain exit $?"; tail -n 1 gpurun_out/plain_fwd.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 100 -c 50 -o /tmp/conv_full $PROF > gpurun_out/ncu_c.log 2>&1; echo "ncu c exit $?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null
python tools/ncu_summarise.py gpurun_out/conv_full_raw.csv gpurun_out/ncu_conv_tc_full.txt gpurun_out/ncu_traffic.json --shape "[16,624,1024,3]"
ls -la gpurun_out/*.csv
