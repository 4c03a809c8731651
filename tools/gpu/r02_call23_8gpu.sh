# Round 2, GPU call 23 (8 GPUs): predict at N = 8 on the final code (engine chunk 16).
set -x
mkdir -p gpurun_out
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $RUN8 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_g8.json') if l.startswith('{')][-1])
print('g8 value %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['clocks'])
PY
