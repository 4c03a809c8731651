# Round 2, GPU call 17 (8 GPUs): predict and train at N = 8 on the final code.
set -x
mkdir -p gpurun_out
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $RUN8 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 exit $?"
timeout 600 $RUN8 bench.py --gpus 8 --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g8.json 2> gpurun_out/bench_train_g8.err; echo "train g8 exit $?"
python - <<'PY'
import json
for f in ('bench_g8', 'bench_train_g8'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['clocks'], d.get('gradient_exchange'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
