set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 exit $?"
tail -n 3 gpurun_out/bench_g2.err; cat gpurun_out/bench_g2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload train > gpurun_out/bench_train_g2.json 2> gpurun_out/bench_train_g2.err; echo "train g2 exit $?"
tail -n 3 gpurun_out/bench_train_g2.err; cat gpurun_out/bench_train_g2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_ref_g2.json 2> gpurun_out/bench_ref_g2.err; echo "ref g2 exit $?"
cat gpurun_out/bench_ref_g2.json
