"""Multi-GPU plumbing: images are independent units, so a batch shards by image with no
collective on the data path; the only exchange is the final gather of per-image mask
canvases to rank 0 (BASELINE.json north_star; the reference itself is single-image,
single-process: /root/reference/serve.py:48).

Host logic only (pure Python + torch.distributed); works with `nccl` on GPUs and with
`gloo` on CPU tensors, which is how tests/ cover world_size 2 without a GPU.
"""
from __future__ import annotations

import numpy as np


def partition_images(costs, world_size):
    """Contiguous partition of images over ranks balancing `costs` (e.g. H*W*N bytes).

    Returns a list of (start, stop) per rank, covering range(len(costs)) in order.
    Contiguity keeps the gathered result in the original image order.  Greedy prefix
    split at multiples of total/world_size.
    """
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    world_size = int(world_size)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if n == 0:
        return [(0, 0)] * world_size
    prefix = np.concatenate([[0.0], np.cumsum(costs)])
    total = prefix[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first index whose prefix reaches the target, but leave enough images for the rest
        i = int(np.searchsorted(prefix, target, side="left"))
        if i > 0 and abs(prefix[i - 1] - target) <= abs(prefix[min(i, n)] - target):
            i -= 1
        i = max(i, bounds[-1])
        i = min(i, n)
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def equal_partition(n_images, world_size):
    """Contiguous blocks of ceil/floor(n/world) images (equal-cost images)."""
    return partition_images(np.ones(n_images), world_size)


def gather_bytes_to_root(local, sizes, dst=0, group=None):
    """Gather ragged uint8 tensors to rank `dst`.

    local: 1-D uint8 tensor of this rank's canvas bytes (device or CPU).
    sizes: list of byte counts per rank (known to all ranks; derived from the partition).
    Returns on dst a list of tensors (one per rank, rank order); None elsewhere.
    Uses point-to-point send/recv (NCCL has no gather primitive; grouped send/recv is what
    torch.distributed.gather lowers to, and it handles ragged sizes).
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if len(sizes) != world:
        raise ValueError("sizes must have one entry per rank")
    if local.numel() != int(sizes[rank]):
        raise ValueError(f"rank {rank}: local has {local.numel()} bytes, sizes says {sizes[rank]}")
    if rank == dst:
        out = []
        ops = []
        for r in range(world):
            if r == dst:
                out.append(local)
            else:
                buf = torch.empty((int(sizes[r]),), dtype=torch.uint8, device=local.device)
                out.append(buf)
                if sizes[r] > 0:
                    ops.append(dist.P2POp(dist.irecv, buf, r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if sizes[rank] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst, group)]):
            req.wait()
    return None
