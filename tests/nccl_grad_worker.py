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
    # (A) the exchange itself, bit for bit: reduce the gradients of a FINISHED backward (its segment events are long
    # recorded) and compare with the sum of the ranks' local copies.  (The weight-gradient kernels accumulate split-K
    # partial sums with f32 atomics, so two backward passes over the same batch differ in the last bits -- the
    # comparison has to use the very buffers that were reduced.)
    tr.forward_backward(imgs, tgt, seed=0)
    torch.cuda.synchronize()
    mine = tr.grads.clone()
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
    # (B) the OVERLAPPED path: the exchange is enqueued right behind the backward, each bucket waiting only for its own
    # segment event.  A bucket reduced before its gradients were final would be off by O(1); what remains is the
    # atomic-order noise of two backward passes.
    tr.forward_backward(imgs, tgt, seed=0, update_stats=False)
    tr.all_reduce_gradients()
    torch.cuda.synchronize()
    err = (tr.grads - total).abs().max().item()
    scale = total.abs().max().item()
    for seg, off, cnt in tr.gradient_buckets():
        a, b = tr.grads[off:off + cnt], total[off:off + cnt]
        rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        assert rel < 1e-3, 'bucket ending at segment %d differs from the finished-backward result by %.3g (relative)' % (seg, rel)
    assert err <= 1e-3 * scale, (err, scale)
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
