"""Data-parallel plumbing for the predict path: images shard trivially, so there is NO collective on the data path.
One process per GPU (``torchrun``); each rank takes a contiguous shard of the dataset order of ``dataset.py:41-68`` and
rank 0 merges the CSV rows back into that order (one ``gather_object`` of small Python lists at the very end).  The
training path's only collective is the gradient all-reduce in ``train.Trainer.optimizer_step``."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init_from_env(device_type='cuda'):
    """Initialise torch.distributed from the torchrun environment (nccl on GPUs, gloo on CPU). Returns (rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if device_type == 'cuda':
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group('gloo')
    return rank, world


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced shard [start, end) of n_items for ``rank`` of ``world``."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def merge_rows(local_rows, start, n_items):
    """Gather each rank's CSV rows (for items [start, start+len)) on rank 0, in dataset order. Returns the merged list on
    rank 0 and None elsewhere; a single process just gets its rows back."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_rows)
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object((start, list(local_rows)), gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = [None] * n_items
    for s, rows in gathered:
        merged[s:s + len(rows)] = rows
    if any(r is None for r in merged):
        raise RuntimeError('merge_rows: shards do not cover the dataset')
    return merged


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device from sysfs (None when it cannot be determined: no sysfs, single-node host, old torch)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = int(p.pci_domain_id), int(p.pci_bus_id), int(p.pci_device_id)
        with open('/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE it allocates pinned host memory:
    first-touch then places the staging buffers on that node and the 50 MB-per-scan host->device copies do not cross the
    socket interconnect.  One process per GPU (torchrun) leaves placement to chance otherwise; with 8 ranks streaming
    raw scans the inter-socket link, not PCIe, becomes the end-to-end bound.  Returns a small report dict."""
    node = gpu_numa_node(device_index)
    info = {'numa_node': node, 'bound': False}
    if node is None:
        return info
    try:
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus))
    except Exception as e:       # containers without sysfs / affinity rights: keep running unbound
        info['error'] = str(e)
    return info
