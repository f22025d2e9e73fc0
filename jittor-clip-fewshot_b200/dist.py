"""Image-sharded data parallelism: one process per GPU, no data-path collective.

The reference is single-process / single-GPU.  Images are independent units (MTA
couples only the views of ONE image, test.py:1692-1742), so the shard unit is the image: rank r owns
the contiguous range [r*I/R, (r+1)*I/R) with all V views of each image.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) carries exactly three tiny control-plane messages:

  broadcast  text embeddings + Channel_LP weights from rank 0     (~2.5 MB, once)
  all_gather per-image top-k predictions int32 [I/R, k]           (KBs, at the end)

Nothing is exchanged between layers, so there is no compute/collective fusion to do on this path.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the job torchrun started (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); single-process otherwise.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def broadcast_tensors(tensors, src=0):
    """In-place broadcast of a list of same-device tensors (text embeddings, head weights)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
    return tensors


def all_gather_topk(local_topk, n_total):
    """Concatenate every rank's [I_r, k] int32 predictions in rank order -> [n_total, k] on every rank.
    Shards may differ by one image, so each is padded to the largest shard before the gather."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local_topk
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    k = local_topk.shape[1]
    pad = torch.zeros((max(sizes), k), dtype=local_topk.dtype, device=local_topk.device)
    pad[: local_topk.shape[0]] = local_topk
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing: the slowest rank defines the step)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
