# Round 2, GPU call 6 (2 GPUs): NCCL gradient-exchange test, predict / train benches at N = 2, bucketed vs single all-reduce.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -x -s -k two_rank > gpurun_out/t_nccl.log 2>&1; echo "nccl test exit $?"
grep -E "NCCL_GRAD_OK|passed|failed|Error|error" gpurun_out/t_nccl.log | head
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $RUN bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 exit $?"
timeout 600 $RUN bench.py --gpus 2 --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g2.json 2> gpurun_out/bench_train_g2.err; echo "train g2 exit $?"
NBC_TRAIN_BUCKETS=0 timeout 600 $RUN bench.py --gpus 2 --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g2_nobucket.json 2> gpurun_out/bench_train_g2_nobucket.err; echo "train g2 nobucket exit $?"
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_g1.json 2> gpurun_out/bench_train_g1.err; echo "train g1 exit $?"
python - <<'PY'
import json
for f in ('bench_g2', 'bench_train_g2', 'bench_train_g2_nobucket', 'bench_train_g1'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['clocks'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 5 gpurun_out/bench_train_g2.err
