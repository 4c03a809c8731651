set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/cli_multi_gpu_check.py 48 > gpurun_out/cli_multi.log 2>&1; echo "cli multi exit $?"
grep -v "Warning\|warn\|^\*\|OMP" gpurun_out/cli_multi.log | tail -n 12
