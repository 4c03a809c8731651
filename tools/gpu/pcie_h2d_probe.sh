set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "ccl or ragged or engine or predict" > gpurun_out/t_ccl.log 2>&1; echo "pytest ccl exit $?"
tail -n 8 gpurun_out/t_ccl.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench_nocpu.json
python - <<'PY'
import torch, time
x = torch.empty(50331648, dtype=torch.uint8).pin_memory()
d = torch.empty(8, 50331648, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(32):
        d[i % 8].copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('pinned H2D %.2f GB/s' % (32 * 50331648 / e0.elapsed_time(e1) / 1e6))
PY
