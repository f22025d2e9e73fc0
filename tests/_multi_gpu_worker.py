"""Worker of tests/test_gpu_multi.py, launched under torch.distributed.run with one rank per GPU: the same seeded
global batch evaluated by ONE rank and by N ranks (uneven speed-balanced shards, both gather modes, host and device
input) must give bit-identical top-5 rows in image order ("1/2/4/8 ranks must produce
bit-identical top-5 to 1 rank"; reference loop test.py:1692-1742, one image at a time)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import jclip_b200 as jb  # noqa: E402


def main():
    rank, world, local = jb.dist.init_from_env()
    assert world >= 2, "launch under torch.distributed.run with >= 2 ranks"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sd = jb.synth.make_vit_state_dict(seed=0)
    model = jb.jclip.build_model(sd)
    largs = types.SimpleNamespace(encoder="vision", position="all", params=["q", "k", "v"], r=4, alpha=1,
                                  dropout_rate=0.25, backbone="ViT-B/32")
    layers = jb.apply_lora(largs, model)
    lora = jb.synth.make_lora(seed=7, b_std=0.05)
    for i, layer in enumerate(layers):
        for name, (A, B) in lora[i].items():
            getattr(layer, name).w_lora_A.data = A
            getattr(layer, name).w_lora_B.data = B
    # rank 0 owns the text banks and the head; the others start from garbage and receive them (the path's broadcast)
    texts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_np = jb.synth.make_head(2, texts[2].numpy())
    bank_t = [(t if rank == 0 else torch.full_like(t, float("nan"))).to(dev) for t in texts]
    lp_dev = [(torch.from_numpy(a) if rank == 0 else torch.zeros(a.shape)).to(dev) for a in lp_np]
    jb.dist.broadcast_tensors(bank_t + lp_dev, src=0)
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = [t.cpu() for t in lp_dev]
    n_total, V = 53, 5
    images = (jb.synth.make_views_torch(77, n_total, V, dev) * 255).round_().to(torch.uint8)   # same on every rank
    for rank_by in ("cs1", "cs5"):
        hp = jb.HotPath(model, jb.TextBank(bank_t[0], bank_t[1], bank_t[2], dev), lp, rank_by=rank_by)
        full = hp.evaluate_base(images, topk_to_host=False)                  # this rank alone, the whole batch
        # every rank computed the same thing on its own GPU: identical across ranks
        ref = full.clone()
        torch.distributed.broadcast(ref, src=0)
        assert torch.equal(ref, full), f"rank {rank}: single-rank result differs from rank 0's"
        for sizes in (jb.dist.shard_sizes(n_total, world),
                      jb.dist.balanced_shard_sizes(n_total, [1.0 + 0.7 * r for r in range(world)])):
            lo, hi = jb.dist.shard_range_from_sizes(sizes, rank)
            mine = hp.evaluate_base(images[lo:hi].contiguous(), topk_to_host=False)
            got = jb.dist.all_gather_topk(mine, n_total, sizes)
            assert torch.equal(got, full), f"rank {rank} sizes {sizes}: gathered top-5 differs from the 1-rank result"
            g = jb.dist.AsyncTopkGather(n_total, 5, dev, sizes=sizes, depth=2)
            tickets = [g.submit(mine) for _ in range(3)]
            assert all(torch.equal(g.result(t), full) for t in tickets[1:])
            g.drain()
            # the end-to-end form: pinned host views in, host top-5 out
            host = images[lo:hi].cpu().pin_memory()
            mine_h = hp.evaluate_base(host)
            assert torch.equal(mine_h, full[lo:hi].cpu())
    jb.dist.barrier()
    if rank == 0:
        print(f"MULTI_OK world={world} images={n_total} views={V}")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
