"""Worker of tests/test_gpu_train.py::test_two_rank_nccl_gradient_exchange (launched with torch.distributed.run, one
process per GPU): the Trainer's bucketed, overlapped NCCL exchange gives exactly the sum of the ranks' local gradients,
the same numbers as one all-reduce of the whole buffer, and identical weights on every rank after the step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import model as omodel, synth
    from neuralbarkcalculator_b200.train import Trainer
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'model_small.npz'))
    sd = omodel.synthetic_state_dict(seed=int(g['state_dict_seed']), head=(g['head_w'], g['head_b']))
    N, H, W = 2, 64, 96
    imgs = torch.from_numpy(np.stack([synth.texture_u8(H, W, 10 * rank + i) for i in range(N)])).to(dev)
    tgt = torch.from_numpy(np.stack([synth.class_mask(H, W, 50 + 10 * rank + i) for i in range(N)])).to(dev)
    tr = Trainer(sd, N, H, W, device=str(dev), dropout=0.0, bucket_mb=16.0)
    assert len(tr.gradient_buckets()) >= 3
    tr.forward_backward(imgs, tgt, seed=0)
    torch.cuda.synchronize()
    mine = tr.grads.clone()
    # one more backward, this time with the exchange enqueued behind it -- the overlapped path
    tr.forward_backward(imgs, tgt, seed=0, update_stats=False)
    n_coll = tr.all_reduce_gradients()
    torch.cuda.synchronize()
    assert n_coll == len(tr.gradient_buckets())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    total = parts[0].clone()
    for p in parts[1:]:
        total += p
    whole = mine.clone()
    dist.all_reduce(whole)
    assert (parts[0] != parts[1]).any(), 'the ranks must see different data'
    if world == 2:      # a + b is exact and commutative: bit-for-bit
        assert torch.equal(tr.grads, total), 'bucketed all-reduce != sum of the local gradients'
        assert torch.equal(tr.grads, whole), 'bucketed != single all-reduce'
    else:               # summation order differs between algorithms
        assert (tr.grads - total).abs().max() <= 1e-5 * total.abs().max()
    # a full step leaves every rank with the same weights (mean gradient: Adam applies 1 / world)
    tr.step(imgs, tgt)
    torch.cuda.synchronize()
    ps = [torch.empty_like(tr.params) for _ in range(world)]
    dist.all_gather(ps, tr.params)
    assert all(torch.equal(ps[0], p) for p in ps[1:]), 'weights diverged across ranks'
    if rank == 0:
        print('NCCL_GRAD_OK world=%d buckets=%d grad_norm=%.6g' % (world, n_coll, float(total.norm())))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
