"""Concurrent pinned host->device bandwidth of a box (run under torchrun, one rank per GPU): every rank copies 50 MB pinned
buffers to its own GPU, (a) one rank at a time, (b) all ranks together, (c) subsets of the ranks -- which GPUs share a host
bridge shows up as pairs whose concurrent rate halves.  usage: torchrun --nproc-per-node N tools/h2d_probe.py"""
import os
import subprocess

import torch
import torch.distributed as dist


def rate(dst, src, reps=24):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dst[i % dst.shape[0]].copy_(src[i % len(src)], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * src[0].numel() / e0.elapsed_time(e1) / 1e6      # GB/s


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n = 50331648
    src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for s in src:
        s.random_(0, 255)
    dst = torch.empty(4, n, dtype=torch.uint8, device='cuda')
    rate(dst, src, 8)
    out = torch.zeros(world, device='cuda')

    def together(active, label):
        dist.barrier()
        r = rate(dst, src) if rank in active else 0.0
        out.zero_()
        out[rank] = r
        dist.all_reduce(out)
        if rank == 0:
            v = out.tolist()
            print('%-26s per GPU [%s] GB/s  sum %.1f' % (label, ' '.join('%5.1f' % x for x in v), sum(v)), flush=True)
    if rank == 0:
        print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout[:2500], flush=True)
        print('host cores', os.cpu_count(), flush=True)
    for r in range(world):
        together({r}, 'alone: GPU %d' % r)
    together(set(range(world)), 'all %d together' % world)
    if world >= 4:
        together(set(range(world // 2)), 'first half')
        together(set(range(world // 2, world)), 'second half')
        together(set(range(0, world, 2)), 'even GPUs')
    if world >= 2:
        for r in range(1, world):
            together({0, r}, 'pair 0 + %d' % r)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
