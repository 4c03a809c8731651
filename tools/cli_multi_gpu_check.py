"""Data-parallel predict.py check (run under torchrun on the GPU box): every rank processes its shard of a synthetic
folder; rank 0 then runs the same folder on one GPU and compares CSV and every PNG byte for byte."""
import argparse
import filecmp
import os
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralbarkcalculator_b200 import distributed as ndist, predict as npredict  # noqa: E402
from oracle import model as omodel, synth  # noqa: E402

rank, local_rank, world = ndist.env_world()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'model_small.npz'))
sd = omodel.synthetic_state_dict(seed=0, head=(g['head_w'], g['head_b']))
root = '/dev/shm/nbc_multi'
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    synth.make_raw_folder(root, n, size=4096, pool=min(n, 6), seed0=900)
ndist.init_from_env('cuda')
torch.cuda.set_device(local_rank)
import torch.distributed as dist  # noqa: E402
dist.barrier()
ns = argparse.Namespace(root_path=root, device='cuda:%d' % local_rank, exclude_nodes=True, only_preprocess=False)
for it in range(2):
    dist.barrier()
    t0 = time.perf_counter()
    rows = npredict.main(ns, state_dict=sd)
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print('world %d: %d images in %.3f s = %.1f img/s (wall, files in, files out)' % (world, n, dt, n / dt))
if rank == 0:
    multi = root + '_multi_out'
    shutil.rmtree(multi, ignore_errors=True)
    os.makedirs(multi)
    shutil.move(os.path.join(root, 'processed'), multi)
    shutil.move(os.path.join(root, 'results'), multi)
    saved = {k: os.environ.pop(k) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}      # single-process run of the same folder
    rows1 = npredict.main(argparse.Namespace(root_path=root, device='cuda:0', exclude_nodes=True, only_preprocess=False), state_dict=sd)
    os.environ.update(saved)
    assert rows == rows1, 'CSV rows differ'
    bad = []
    for top in ('processed', 'results'):
        for d, _, files in os.walk(os.path.join(root, top)):
            for f in files:
                a = os.path.join(d, f)
                b = os.path.join(multi, os.path.relpath(a, root))
                if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                    bad.append(a)
    assert not bad, bad[:5]
    print('multi-GPU outputs identical to the single-GPU run: %d rows, all files byte-identical' % (len(rows) - 1))
    shutil.rmtree(root, ignore_errors=True)
    shutil.rmtree(multi, ignore_errors=True)
dist.barrier()
dist.destroy_process_group()
