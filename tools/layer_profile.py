"""Per-layer CUDA-event profile of the native forward (run on the GPU box).
usage: python tools/layer_profile.py [N] [H] [W] [impl]   -> table: layer, shape, ms, TFLOP/s, GB/s (algorithmic)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralbarkcalculator_b200 as nbc  # noqa: E402
from oracle import model as omodel, synth  # noqa: E402


def layer_names(fused=True):
    names = ['stem 3->64 k7 s2', 'maxpool']
    inpl = 64
    for li, (nb, planes) in enumerate(zip((3, 4, 6, 3), (64, 128, 256, 512))):
        for b in range(nb):
            pre = 'layer%d.%d.' % (li + 1, b)
            names.append(pre + 'conv1 %d->%d k1' % (inpl, planes))
            names.append(pre + 'conv2 %d->%d k3' % (planes, planes))
            if b == 0 and not fused:
                names.append(pre + 'downsample %d->%d k1' % (inpl, planes * 4))
            names.append(pre + ('conv3+downsample %d+%d->%d' % (planes, inpl, planes * 4) if (b == 0 and fused) else
                                'conv3 %d->%d k1' % (planes, planes * 4)))
            inpl = planes * 4
    names += ['head conv3x3 2048->512', 'head 1x1 512->3']
    return names


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'model_small.npz'))
    sd = omodel.synthetic_state_dict(seed=0, head=(g['head_w'], g['head_b']))
    dev = torch.device('cuda:0')
    m = nbc.fcn_resnet50(pretrained=False)
    m.load_state_dict(sd)
    m.to(dev).eval()
    m.set_normalisation(omodel.DEFAULT_MEAN, omodel.DEFAULT_STD)
    plan = m.native_plan()
    plan.set_impl(impl)
    img = torch.from_numpy(synth.texture_u8(H, W, 1)).to(dev).unsqueeze(0).repeat(N, 1, 1, 1).contiguous()
    for _ in range(3):
        plan.profile(img)
    runs = [plan.profile(img) for _ in range(5)]
    ms = np.median(np.array([[r[0] for r in run] for run in runs]), axis=0)
    fl = [r[1] for r in runs[0]]
    names = layer_names(fused=(impl != 2))
    assert len(names) == len(ms), (len(names), len(ms))
    print('N=%d H=%d W=%d impl=%d' % (N, H, W, impl))
    tot = 0.0
    for n, t, f in zip(names, ms, fl):
        print('%-38s %8.3f ms  %8.1f TFLOP/s' % (n, t, f / (t * 1e-3) / 1e12 if t > 0 else 0))
        tot += t
    conv = [(t, f) for n, t, f in zip(names, ms, fl) if 'conv' in n or 'downsample' in n]
    ct, cf = sum(t for t, _ in conv), sum(f for _, f in conv)
    print('TOTAL %.3f ms (%.3f ms/img); tensor-core convs %.3f ms, %.1f TFLOP/s; %.1f img/s' % (tot, tot / N, ct, cf / (ct * 1e-3) / 1e12, N / (tot * 1e-3)))


if __name__ == '__main__':
    main()
