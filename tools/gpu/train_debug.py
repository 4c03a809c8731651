"""Layer-by-layer comparison of the native train-mode forward with the torch oracle (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as omodel, synth, losses as olosses
from neuralbarkcalculator_b200.train import Trainer

g = np.load('tests/golden/model_small.npz')
sd = omodel.synthetic_state_dict(0, head=(g['head_w'], g['head_b']))
N, H, W = 2, 64, 96
imgs = np.stack([synth.texture_u8(H, W, i) for i in range(N)])
tgt = np.stack([synth.class_mask(H, W, 100 + i) for i in range(N)])
x = torch.stack([omodel.normalise_u8(im)[0] for im in imgs])
net = omodel.fcn_resnet50(dropout=0.0); net.load_state_dict(sd); net.train()
acts = {}
def hook(name):
    def f(m, i, o): acts[name] = o.detach()
    return f
names = []
for name, m in net.named_modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
        m.register_forward_hook(hook(name)); names.append(name)
logits = net(x)
loss = olosses.custom_weighted_cross_entropy(logits, torch.from_numpy(tgt).long(), torch.ones(3))
tr = Trainer(sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=torch.ones(3))
l = tr.forward_backward(torch.from_numpy(imgs).cuda(), torch.from_numpy(tgt).cuda(), seed=1)
print('loss ours %.5f oracle %.5f' % (float(l), float(loss)))
keys = [k for k in sd.keys() if k.endswith('.weight') and sd[k].dim() == 4][:-1]   # conv weights in state_dict order
for ui, k in enumerate(keys):
    conv_name = k[:-len('.weight')]
    z = tr.debug_tensor(0, ui).float().cpu().permute(0, 3, 1, 2)
    zr = acts[conv_name]
    err = (z - zr).abs()
    print('%-36s z: ref std %.4f mean %.4f | err max %.4f mean %.5f (rel %.4f)' % (conv_name, zr.std(), zr.mean(), err.max(), err.mean(), err.mean() / zr.std()))
low = tr.debug_tensor(2).cpu(); full = tr.debug_tensor(3).cpu()
print('full logits err max %.4f mean %.5f ref std %.3f' % ((full - logits.detach()).abs().max(), (full - logits.detach()).abs().mean(), logits.std()))
