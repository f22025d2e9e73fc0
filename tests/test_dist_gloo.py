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
    assert jb.dist.max_over_ranks(2.5, "cpu") == 2.5
    jb.dist.barrier()
