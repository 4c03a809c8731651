set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -k "lovasz or metrics or mixed" -s > gpurun_out/t_lovasz.log 2>&1; echo "pytest exit $?"
grep -v "Warning\|warn" gpurun_out/t_lovasz.log | tail -n 40
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from neuralbarkcalculator_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
for (N, H, W) in ((5, 512, 512), (8, 1024, 1024)):
    logits = torch.randn(N, 3, H, W, device='cuda', generator=g)
    target = torch.randint(0, 3, (N, H, W), device='cuda', generator=g, dtype=torch.uint8)
    for _ in range(2):
        ops.lovasz_softmax_fwd_bwd(logits, target)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss, grad = ops.lovasz_softmax_fwd_bwd(logits, target)
    e1.record(); torch.cuda.synchronize()
    print('lovasz fwd+bwd N=%d %dx%d: %.3f ms (%.1f Mpx/s)' % (N, H, W, e0.elapsed_time(e1) / 5, N * H * W * 5 / e0.elapsed_time(e1) / 1e3))
PY
