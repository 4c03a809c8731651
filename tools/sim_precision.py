"""CPU simulation of the storage-precision error budget of the native plan (test / design tooling, not product).

Replays the FCN-ResNet50 forward the way plan.cu runs it -- BN folded into the weights, weights and activations rounded
to a 16-bit format at the points where the GPU stores them, f32 accumulation, residual added in f32 before the output is
rounded -- and compares with the f32 oracle.  Variants switch single error sources off so the budget can be read:

    python tools/sim_precision.py [H W]

Only ``oracle`` (test infrastructure) and torch are used."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import model as omodel   # noqa: E402
from oracle import synth             # noqa: E402


def q(x, dt):
    return x if dt is None else x.to(dt).float()


def fold(sd, conv, bn):
    w = sd[conv + '.weight'].double()
    g, b = sd[bn + '.weight'].double(), sd[bn + '.bias'].double()
    m, v = sd[bn + '.running_mean'].double(), sd[bn + '.running_var'].double()
    s = g / torch.sqrt(v + 1e-5)
    return (w * s.view(-1, 1, 1, 1)).float(), (b - m * s).float()


@torch.no_grad()
def forward_sim(sd, x, wdt, adt, trunk_dt='same', head_dt='same', trunk_split=False):
    """wdt: weight storage dtype; adt: inner activation dtype; trunk_dt: block-output dtype ('same' = adt);
    trunk_split: the trunk is stored as hi (adt, feeds the convs) + lo (adt, residual only)."""
    if trunk_dt == 'same':
        trunk_dt = adt
    if head_dt == 'same':
        head_dt = adt

    def conv(x, name, bn, stride=1, pad=0, dil=1, relu=True, out_dt=adt, extra=None):
        w, b = fold(sd, name, bn)
        y = F.conv2d(x, q(w, wdt), b, stride=stride, padding=pad, dilation=dil)
        if extra is not None:
            y = y + extra
        if relu:
            y = F.relu(y)
        return q(y, out_dt)

    x = q(x, adt)
    y = conv(x, 'backbone.conv1', 'backbone.bn1', stride=2, pad=3)
    y = F.max_pool2d(y, 3, 2, 1)
    trunk_exact = y      # what the residual add sees
    trunk = y            # what the convs see
    dil = 1
    cfg = [(1, 3, 1, False), (2, 4, 2, False), (3, 6, 1, True), (4, 3, 1, True)]
    for li, nblocks, stride, dilate in cfg:
        for bi in range(nblocks):
            p = 'backbone.layer%d.%d.' % (li, bi)
            s = stride if (bi == 0 and not dilate) else 1
            prev_dil = dil
            if bi == 0 and dilate:
                dil = dil * (2 if li >= 3 else 1)
            d2 = prev_dil if bi == 0 else dil
            a = conv(trunk, p + 'conv1', p + 'bn1')
            a = conv(a, p + 'conv2', p + 'bn2', stride=s, pad=d2, dil=d2)
            if bi == 0:
                wd, bd = fold(sd, p + 'downsample.0', p + 'downsample.1')
                idn = F.conv2d(trunk, q(wd, wdt), bd, stride=s)      # fused into conv3 as extra K: f32, never stored
            else:
                idn = trunk_exact
            w3, b3 = fold(sd, p + 'conv3', p + 'bn3')
            v = F.relu(F.conv2d(a, q(w3, wdt), b3) + idn)
            if trunk_split:
                trunk = q(v, trunk_dt)
                trunk_exact = trunk + q(v - trunk, trunk_dt)
            else:
                trunk = trunk_exact = q(v, trunk_dt)
    h = conv(trunk, 'classifier.0', 'classifier.1', pad=1, out_dt=head_dt)
    return F.conv2d(h, sd['classifier.4.weight'], sd['classifier.4.bias'])


def report(name, got, ref, size):
    err = (got - ref).abs()
    up_r = F.interpolate(ref, size=size, mode='bicubic', align_corners=False)
    up_g = F.interpolate(got, size=size, mode='bicubic', align_corners=False)
    agree = (up_r.argmax(1) == up_g.argmax(1)).float().mean().item()
    pr = [(up_r.argmax(1) == c).float().mean().item() * 100 for c in (1, 2)]
    pg = [(up_g.argmax(1) == c).float().mean().item() * 100 for c in (1, 2)]
    print('%-44s max-abs %.5f mean-abs %.5f (std %.3f) full-res max %.5f agree %.5f  pp %.4f %.4f'
          % (name, err.max(), err.mean(), ref.std(), (up_g - up_r).abs().max(), agree, abs(pr[0] - pg[0]), abs(pr[1] - pg[1])))


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 256)
    torch.set_num_threads(os.cpu_count())
    h16, b16 = torch.float16, torch.bfloat16
    for label, kw in (('harsh (whiten, gain 0.5)', dict()),
                      ('trained-like (features, gain 0.1)', dict(calibration='features', branch_gain=0.1)),
                      ('features, gain 0.5', dict(calibration='features', branch_gain=0.5))):
        sd = omodel.synthetic_state_dict(seed=0, **kw)
        net = omodel.load_model(sd)
        img = synth.texture_u8(H, W, 21)
        x = omodel.normalise_u8(img)
        with torch.no_grad():
            ref = net.features(x)
        print('== %s, %dx%d' % (label, H, W))
        report('f32 restatement (fold only)', forward_sim(sd, x, None, None), ref, (H, W))
        report('bf16 all', forward_sim(sd, x, b16, b16), ref, (H, W))
        report('fp16 all', forward_sim(sd, x, h16, h16), ref, (H, W))
        report('fp16 weights only', forward_sim(sd, x, h16, None), ref, (H, W))
        report('fp16 activations only', forward_sim(sd, x, None, h16), ref, (H, W))
        report('fp16 inner acts only (trunk f32)', forward_sim(sd, x, None, h16, trunk_dt=None, head_dt=None), ref, (H, W))
        report('fp16 trunk only', forward_sim(sd, x, None, None, trunk_dt=h16, head_dt=None), ref, (H, W))
        report('fp16 all, trunk split hi+lo', forward_sim(sd, x, h16, h16, trunk_split=True), ref, (H, W))
        report('fp16 all, trunk split, head f32', forward_sim(sd, x, h16, h16, trunk_split=True, head_dt=None), ref, (H, W))
        report('fp16 all, head f32', forward_sim(sd, x, h16, h16, head_dt=None), ref, (H, W))


if __name__ == '__main__':
    main()
