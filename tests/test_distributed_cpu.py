"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding + ordered CSV merge give the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralbarkcalculator_b200 import distributed as nd


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    r, w = nd.init_from_env(device_type='cpu')
    assert (r, w) == (rank, world)
    start, end = nd.shard_bounds(n_items, rank, world)
    rows = [['img_%04d.png' % i, 'sapin', '%.5f' % (i * 1.5)] for i in range(start, end)]     # stand-in for per-image stats
    merged = nd.merge_rows(rows, start, n_items)
    if rank == 0:
        torch.save(merged, out_path)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [7, 2, 64])
def test_two_rank_merge_equals_single_process(tmp_path, n_items):
    out = str(tmp_path / 'merged.pt')
    mp.spawn(_worker, args=(2, _free_port(), n_items, out), nprocs=2, join=True)
    merged = torch.load(out)
    expect = [['img_%04d.png' % i, 'sapin', '%.5f' % (i * 1.5)] for i in range(n_items)]
    assert merged == expect


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 64, 10001):
        for world in (1, 2, 4, 8):
            b = [nd.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    assert nd.merge_rows([[1], [2]], 0, 2) == [[1], [2]]     # single process: passthrough


def _bucket_worker(rank, world, port, out_path):
    """The Trainer's bucketed gradient exchange on CPU tensors over gloo: real segment table of the native plan, one async
    all-reduce per bucket, against one all-reduce of the whole flat buffer."""
    import ctypes as C
    from neuralbarkcalculator_b200 import _lib
    from neuralbarkcalculator_b200.train import gradient_segments, merge_segments
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    nd.init_from_env(device_type='cpu')
    lib = _lib.load()
    h = lib.nbc_train_create(1, 64, 64)          # host-side bookkeeping only: no CUDA call
    n = lib.nbc_train_param_count(C.c_void_p(h))
    segs = gradient_segments(lib, h)
    buckets = merge_segments(segs, 16e6)
    lib.nbc_train_destroy(C.c_void_p(h))
    g = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(n, generator=g)
    whole = grads.clone()
    dist.all_reduce(whole)
    works = [dist.all_reduce(grads[off:off + cnt], async_op=True) for _, off, cnt in buckets]
    for w in works:
        w.wait()
    assert torch.equal(grads, whole)
    if rank == 0:
        torch.save({'n': n, 'segments': segs, 'buckets': buckets}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_two_ranks(tmp_path):
    out = str(tmp_path / 'buckets.pt')
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = torch.load(out)
    segs, buckets = d['segments'], d['buckets']
    assert d["n"] == 32947779 and len(segs) == 18          # 32 947 779 parameters, 16 blocks + head + stem
    assert segs[0][0] + segs[0][1] == d['n'] and segs[-1][0] == 0
    assert all(segs[i][0] == segs[i + 1][0] + segs[i + 1][1] for i in range(len(segs) - 1))
    assert sum(c for _, _, c in buckets) == d['n'] and 3 <= len(buckets) <= 8
    assert all(c * 4 >= 16e6 for _, _, c in buckets[:-1])
