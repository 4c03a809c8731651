set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 exit $?"
tail -n 3 gpurun_out/bench_g8.err; cat gpurun_out/bench_g8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --workload train > gpurun_out/bench_train_g8.json 2> gpurun_out/bench_train_g8.err; echo "train g8 exit $?"
tail -n 3 gpurun_out/bench_train_g8.err; cat gpurun_out/bench_train_g8.json
