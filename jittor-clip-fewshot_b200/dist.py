"""Image-sharded data parallelism: one process per GPU, no data-path collective.

The reference is single-process / single-GPU.  Images are independent units (MTA
couples only the views of ONE image, test.py:1692-1742), so the shard unit is the image: rank r owns
the contiguous range [r*I/R, (r+1)*I/R) with all V views of each image.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) carries exactly three tiny control-plane messages:

  broadcast  text embeddings + Channel_LP weights from rank 0     (~2.5 MB, once)
  all_gather per-image top-k predictions int32 [I_r, k]           (KBs, per step)

Shard sizes are equal by default; `balanced_shard_sizes` sizes them by each rank's measured speed (the GPUs of one
box differ by a few per cent under the power cap, and the per-step all-gather waits for the slowest).

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


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs closest to its GPU (NVML's ideal CPU affinity = the GPU's NUMA node), so that the
    pinned staging buffers it allocates afterwards are first-touched on that node and host->device copies do not cross
    the socket interconnect.  Best effort: returns the CPU count bound to, or None when NVML / the call is unavailable."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(int(local_rank)))
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None


def shard_range(n_items, rank, world):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def balanced_shard_sizes(n_items, seconds_per_item):
    """Shard sizes proportional to each rank's measured speed (1 / seconds per item), summing to n_items.

    The GPUs of one box do not run at the same speed under the power cap (measured spread of the same step: a few per
    cent), and with equal shards every rank waits for the slowest one at the per-step all-gather.  Images are independent,
    so the shard boundaries are free: give each rank work in proportion to what it gets through per second.  Largest-
    remainder rounding; every rank keeps at least one item when n_items >= world."""
    speed = [1.0 / max(float(t), 1e-12) for t in seconds_per_item]
    world = len(speed)
    n_items = int(n_items)
    total = sum(speed)
    exact = [n_items * s / total for s in speed]
    sizes = [int(e) for e in exact]
    if n_items >= world:
        sizes = [max(s, 1) for s in sizes]
    # hand out (or take back) what rounding left, largest fractional part first
    order = sorted(range(world), key=lambda r: exact[r] - int(exact[r]), reverse=True)
    i = 0
    while sum(sizes) < n_items:
        sizes[order[i % world]] += 1
        i += 1
    i = 0
    while sum(sizes) > n_items:
        r = order[::-1][i % world]
        if sizes[r] > (1 if n_items >= world else 0):
            sizes[r] -= 1
        i += 1
    return sizes


def shard_range_from_sizes(sizes, rank):
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def all_gather_floats(value, device):
    """Every rank's python float, in rank order (calibration timings)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def broadcast_tensors(tensors, src=0):
    """In-place broadcast of a list of same-device tensors (text embeddings, head weights)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
    return tensors


def all_gather_topk(local_topk, n_total, sizes=None):
    """Concatenate every rank's [I_r, k] int32 predictions in rank order -> [n_total, k] on every rank.
    Shards may differ in size (by one image for the even split, by more for `balanced_shard_sizes`), so each is padded
    to the largest shard before the gather.  `sizes`: the per-rank shard sizes if not the even split."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local_topk
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world) if sizes is None else [int(x) for x in sizes]
    if len(sizes) != world or sum(sizes) != n_total or local_topk.shape[0] != sizes[dist.get_rank()]:
        raise ValueError(f"shard sizes {sizes} do not describe {n_total} items over {world} ranks "
                         f"(this rank holds {local_topk.shape[0]})")
    k = local_topk.shape[1]
    pad = torch.zeros((max(sizes), k), dtype=local_topk.dtype, device=local_topk.device)
    pad[: local_topk.shape[0]] = local_topk
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


class AsyncTopkGather:
    """The per-step all-gather of the top-k predictions, taken off the compute stream's critical path.

    `all_gather_topk` makes the stream that computes step k+1 wait for the gather of step k, i.e. for the slowest rank,
    every step.  Nothing in step k+1 depends on it, so here the gather is launched asynchronously (NCCL's own stream waits
    for the predictions, the compute stream does not wait for NCCL) into one of `depth` persistent slots; `result(ticket)`
    / `drain()` make the current stream wait when the predictions are actually consumed.  Ranks then run at their own
    pace, at most `depth` steps apart.  Single process: a pass-through."""

    def __init__(self, n_total, k, device, sizes=None, depth=4, dtype=torch.int32):
        self.multi = dist.is_initialized() and dist.get_world_size() > 1
        self.n_total, self.k, self.depth = int(n_total), int(k), max(int(depth), 1)
        self.world = dist.get_world_size() if self.multi else 1
        self.rank = dist.get_rank() if self.multi else 0
        self.sizes = shard_sizes(self.n_total, self.world) if sizes is None else [int(x) for x in sizes]
        if len(self.sizes) != self.world or sum(self.sizes) != self.n_total:
            raise ValueError(f"shard sizes {self.sizes} do not describe {n_total} items over {self.world} ranks")
        self.turn = 0
        self.local = [None] * self.depth
        self.work = [None] * self.depth
        if self.multi:
            m = max(self.sizes)
            self.pad = [torch.zeros((m, self.k), dtype=dtype, device=device) for _ in range(self.depth)]
            self.out = [torch.empty((self.world * m, self.k), dtype=dtype, device=device) for _ in range(self.depth)]
            self.into_tensor = dist.get_backend() == "nccl"

    def submit(self, local_topk):
        """Enqueue the gather of this rank's [I_r, k] predictions; returns a ticket for `result`."""
        t = self.turn % self.depth
        self.turn += 1
        if not self.multi:
            self.local[t] = local_topk
            return t
        if local_topk.shape[0] != self.sizes[self.rank]:
            raise ValueError(f"this rank's shard has {self.sizes[self.rank]} items, got {local_topk.shape[0]}")
        if self.work[t] is not None:
            self.work[t].wait()                  # the slot's previous gather (depth steps back) has read `pad`
        self.pad[t][: local_topk.shape[0]].copy_(local_topk, non_blocking=True)
        if self.into_tensor:
            self.work[t] = dist.all_gather_into_tensor(self.out[t], self.pad[t], async_op=True)
        else:                                    # gloo (CPU tests): list form
            m = self.pad[t].shape[0]
            self.work[t] = dist.all_gather([self.out[t][r * m:(r + 1) * m] for r in range(self.world)], self.pad[t], async_op=True)
        return t

    def result(self, ticket):
        """[n_total, k] predictions of the step that returned `ticket` (valid until the slot is reused, depth steps on)."""
        if not self.multi:
            return self.local[ticket]
        if self.work[ticket] is not None:
            self.work[ticket].wait()
        m = self.pad[ticket].shape[0]
        return torch.cat([self.out[ticket][r * m: r * m + s] for r, s in enumerate(self.sizes)], dim=0)

    def drain(self):
        """Make the current stream wait for every gather still in flight (end of a timed region)."""
        if self.multi:
            for w in self.work:
                if w is not None:
                    w.wait()


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
