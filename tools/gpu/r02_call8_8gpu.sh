# Round 2, GPU call 8 (8 GPUs): predict / train at N = 8 (and 4), concurrent pinned-H2D probe, NCCL gradient test.
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
RUN4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532"
timeout 300 $RUN8 tools/h2d_probe.py > gpurun_out/h2d_probe_8gpu.txt 2> gpurun_out/h2d_probe_8gpu.err; echo "probe exit $?"; grep -E "GB/s|cores" gpurun_out/h2d_probe_8gpu.txt
timeout 600 $RUN8 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 exit $?"
timeout 600 $RUN4 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "bench g4 exit $?"
NBC_ZERO_SPAN=0 timeout 600 $RUN8 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_g8_nospan.json 2> gpurun_out/bench_g8_nospan.err; echo "bench g8 nospan exit $?"
timeout 600 $RUN8 bench.py --gpus 8 --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g8.json 2> gpurun_out/bench_train_g8.err; echo "train g8 exit $?"
NBC_TRAIN_BUCKETS=0 timeout 600 $RUN8 bench.py --gpus 8 --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g8_nobucket.json 2> gpurun_out/bench_train_g8_nobucket.err; echo "train g8 nobucket exit $?"
timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -x -s -k two_rank > gpurun_out/t_nccl.log 2>&1; echo "nccl test exit $?"
grep -E "NCCL_GRAD_OK|passed|failed|Assertion" gpurun_out/t_nccl.log | head
python - <<'PY'
import json
for f in ('bench_g8', 'bench_g4', 'bench_g8_nospan', 'bench_train_g8', 'bench_train_g8_nobucket'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['clocks'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
