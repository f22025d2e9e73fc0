"""CPU, world_size 2 over gloo: the N>1 host logic of the image-sharded path -- shard ranges, the
broadcast of text embeddings / head weights, the all-gather of per-image top-5 with uneven shards,
and the max-over-ranks timing reduction."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import jclip_b200 as jb
    r, w, _ = jb.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # rank 0 owns the text bank / head; everyone else starts from garbage
    text = torch.arange(403 * 8, dtype=torch.float32).view(403, 8) if rank == 0 else torch.full((403, 8), -1.0)
    head = torch.ones(8) * 3 if rank == 0 else torch.zeros(8)
    jb.dist.broadcast_tensors([text, head], src=0)
    assert text[5, 3] == 5 * 8 + 3 and head[0] == 3
    lo, hi = jb.dist.shard_range(n_images, rank, world)
    # each rank "predicts" top-5 = image index + j for its own shard
    local = (torch.arange(lo, hi, dtype=torch.int32).view(-1, 1) + torch.arange(5, dtype=torch.int32).view(1, -1))
    full = jb.dist.all_gather_topk(local, n_images)
    assert full.shape == (n_images, 5)
    assert torch.equal(full[:, 0], torch.arange(n_images, dtype=torch.int32))
    # speed-balanced shards: rank 0 is "twice as fast", so it takes about two thirds of the images
    times = jb.dist.all_gather_floats(1.0 + rank, "cpu")
    assert times == [1.0, 2.0]
    sizes = jb.dist.balanced_shard_sizes(n_images, times)
    assert sum(sizes) == n_images and sizes[0] > sizes[1] >= 1
    lo2, hi2 = jb.dist.shard_range_from_sizes(sizes, rank)
    local2 = (torch.arange(lo2, hi2, dtype=torch.int32).view(-1, 1) + torch.arange(5, dtype=torch.int32).view(1, -1))
    full2 = jb.dist.all_gather_topk(local2, n_images, sizes)
    assert torch.equal(full2, full)
    # the asynchronous form: several gathers in flight, slots reused
    g = jb.dist.AsyncTopkGather(n_images, 5, "cpu", sizes=sizes, depth=2)
    tickets = [g.submit(local2 + step) for step in range(3)]
    assert torch.equal(g.result(tickets[2]), full + 2) and torch.equal(g.result(tickets[1]), full + 1)
    g.drain()
    with pytest.raises(ValueError):
        jb.dist.all_gather_topk(local2, n_images, [n_images, 0] if sizes != [n_images, 0] else [0, n_images])
    assert jb.dist.max_over_ranks(10.0 + rank, "cpu") == 10.0 + world - 1
    assert jb.dist.sum_over_ranks(1.0, "cpu") == world
    jb.dist.barrier()
    ret[rank] = full.sum().item()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 16])
def test_two_rank_sharding_and_gather(n_images):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] == ret[1]


def test_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    import jclip_b200 as jb
    t = torch.ones(3, 5, dtype=torch.int32)
    assert jb.dist.all_gather_topk(t, 3) is t
    g = jb.dist.AsyncTopkGather(3, 5, "cpu")
    assert g.result(g.submit(t)) is t
    g.drain()
    assert jb.dist.max_over_ranks(2.5, "cpu") == 2.5
    jb.dist.barrier()


def test_balanced_shard_sizes():
    sys.path.insert(0, ROOT)
    import jclip_b200 as jb
    f = jb.dist.balanced_shard_sizes
    assert f(1024, [1.0] * 8) == [128] * 8
    s = f(1024, [75, 76, 77, 80, 75, 74, 79, 78])
    assert sum(s) == 1024 and max(s) == s[5] and min(s) == s[3] and max(s) - min(s) <= 12
    # the predicted step time (size x seconds per item) is flatter than with equal shards
    t = [75, 76, 77, 80, 75, 74, 79, 78]
    assert max(a * b for a, b in zip(s, t)) < 128 * max(t)
    assert f(7, [1, 1]) in ([4, 3], [3, 4]) and f(2, [1, 100]) == [1, 1] and sum(f(1, [1, 2, 3])) == 1 and f(0, [1, 2]) == [0, 0]
    assert jb.dist.shard_range_from_sizes([3, 4, 5], 1) == (3, 7)
