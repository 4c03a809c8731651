set -x
mkdir -p gpurun_out
lscpu | grep -E "NUMA|Socket|Model name" | head -6
nvidia-smi topo -m 2>/dev/null | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 exit $?"
tail -n 2 gpurun_out/bench_g8.err; cut -c1-1300 gpurun_out/bench_g8.json
