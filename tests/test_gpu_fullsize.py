"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle cannot run these in
seconds): config 2 = batch 1024 views through the LoRA tower; config 5 = thousands of images x N in {1, 16, 64}
crops through the whole pipeline.  Properties: a sub-sample matches the fp32 oracle, results do not depend on
how the batch is split into passes (bit-exact), views / images are processed independently (permutation
equivariance, bit-exact), and a batched call equals per-image calls."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double(), b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))


@pytest.fixture(scope="module")
def lora_model(jb):
    sd = jb.synth.make_vit_state_dict(seed=0)
    model = jb.jclip.build_model(sd)
    args = types.SimpleNamespace(encoder="vision", position="all", params=["q", "k", "v"], r=4, alpha=1,
                                 dropout_rate=0.25, backbone="ViT-B/32")
    layers = jb.apply_lora(args, model)
    lora = jb.synth.make_lora(seed=7, b_std=0.05)
    for i, layer in enumerate(layers):
        for name, (A, B) in lora[i].items():
            getattr(layer, name).w_lora_A.data = A
            getattr(layer, name).w_lora_B.data = B
    return sd, lora, model


def test_batch_1024_encode_image(jb, cuda_dev, lora_model):
    """BASELINE config 2: ViT-B/32 + LoRA encode_image, batch 1024."""
    from oracle import vit_encode_image
    sd, lora, model = lora_model
    x = jb.synth.make_views_torch(5, 16, 64, cuda_dev).reshape(1024, 3, 224, 224)
    ctx = jb.get_context(cuda_dev)
    f = model.visual(x, apply_clip_norm=True, normalize=True)
    assert f.shape == (1024, 512) and torch.isfinite(f).all()
    assert (f.norm(dim=-1) - 1).abs().max() < 1e-5
    # a sub-sample against the fp32 oracle
    idx = torch.tensor([0, 1, 63, 64, 500, 777, 1022, 1023])
    ref = vit_encode_image(sd, x[idx.to(cuda_dev)].cpu().numpy(), lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True)
    assert _cos(f[idx.to(cuda_dev)].cpu(), ref).min() >= 0.9995
    # pass-size invariance: 1 pass of 1024 == 4 passes of 256 == 7 ragged passes, bit for bit
    try:
        for bound in (256, 150):
            ctx.set_chunk_views(bound)
            assert torch.equal(model.visual(x, apply_clip_norm=True, normalize=True), f), bound
    finally:
        ctx.set_chunk_views(16384)
    # views are independent: permuting the batch permutes the embeddings, bit for bit
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(0)).to(cuda_dev)
    assert torch.equal(model.visual(x[perm].contiguous(), apply_clip_norm=True, normalize=True), f[perm])
    # determinism
    assert torch.equal(model.visual(x, apply_clip_norm=True, normalize=True), f)


@pytest.mark.parametrize("n_crops,n_images", [(1, 2048), (16, 512), (64, 192)])
def test_pipeline_sweep_sizes(jb, cuda_dev, lora_model, n_crops, n_images):
    """BASELINE config 5 (scaled to one GPU's minute budget): images x N in {1, 16, 64} crops."""
    from oracle import pipeline_image
    sd, lora, model = lora_model
    V = n_crops + 1
    imgs = (jb.synth.make_views_torch(11, n_images, V, cuda_dev) * 255).round_().to(torch.uint8)
    Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_np = jb.synth.make_head(2, Ts[2].numpy())
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = lp_np
    hp = jb.HotPath(model, jb.TextBank(Ts[0], Ts[1], Ts[2], cuda_dev), lp, rank_by="cs5")
    topk, feats, scores = hp.evaluate_base(imgs, return_feats=True, return_scores=True)
    assert topk.shape == (n_images, 5) and int(topk.min()) >= 0 and int(topk.max()) < 403
    assert all(len(set(r)) == 5 for r in topk.cpu().tolist())
    # images are independent: a permutation of the images permutes the predictions, bit for bit
    perm = torch.randperm(n_images, generator=torch.Generator().manual_seed(1)).to(cuda_dev)
    assert torch.equal(hp.evaluate_base(imgs[perm].contiguous()), topk[perm])
    # batched == one image at a time (a handful), and the oracle agrees on those given the same embeddings
    lp_t = tuple(torch.from_numpy(a) for a in lp_np)
    for i in (0, n_images // 2, n_images - 1):
        single = hp.evaluate_base(imgs[i:i + 1].contiguous())
        assert torch.equal(single[0], topk[i])
        fi = feats[i].cpu()
        t5, sc, _ = pipeline_image(fi, fi, Ts[0], Ts[1], Ts[2], lp_t, score="cs5")
        assert (scores[i].cpu() - sc["cs5"][0]).abs().max() <= 1e-2
        assert set(t5.tolist()) == set(topk[i].cpu().tolist())


# End-to-end acceptance (BASELINE.json north star): embedding cosine >= 0.999, logits within 1e-2 absolute, >= 99.5 %
# top-5 label agreement against the fp32 reference on identical synthetic inputs (random unit text
# rows).  Asserted per 16-bit operand type:
#   f16  (default)  the north-star numbers themselves
#   bf16            within 1.2 x of what 8-bit significands measure (0.081 logits, 99.4 % / 97.5 % labels;
#                   profiles/r02_e2e_agreement_*.json and
#                   profiles/r02_error_attribution_*.json: the same deviations reproduced on the CPU by rounding the
#                   fp32 oracle's operands -- half of it is weight rounding, identical for every view, which MTA's
#                   averaging over views cannot remove)
# "structured" text banks (synth.make_structured_text_banks) point the class rows along the image-specific directions
# of the embeddings -- the worst case for the logit deviation (|dlogit| <= 100 |df|) and the case where the top scores are
# whole logits apart; there the gate is the top-5 agreement.  The logit deviation is recorded but not bounded: with
# text rows along the directions in which the views of an image differ, solve_mta's mode can settle on a different
# cluster of views after a 1e-3 perturbation (measured: 0.27 logits with fp16 operands, 1.7 with bf16, on 1 of 64
# images; the fp32 oracle perturbed by the same amount on the CPU does the same, profiles/r02_error_attribution_structured_*).
E2E_BOUNDS = {
    ("f16", "random"): {"dlogit": 1e-2, "cs1": 0.995, "cs5": 0.995},
    ("bf16", "random"): {"dlogit": 0.10, "cs1": 0.99, "cs5": 0.97},
    ("f16", "structured"): {"dlogit": None, "cs1": 0.995, "cs5": 0.995},
    ("bf16", "structured"): {"dlogit": None, "cs1": 0.985, "cs5": 0.985},
}
_oracle_feats_cache = {}


def _oracle_feats(jb, sd, lora, I, V):
    """fp32 oracle tower on the test's views, once per (I, V): [I, V, 512] unit rows (CPU, all host cores)."""
    import os
    from oracle import vit_encode_image
    key = (I, V)
    if key not in _oracle_feats_cache:
        torch.set_num_threads(os.cpu_count() or 1)
        imgs = jb.synth.make_views(21, I, V)                       # float32 [I, V, 3, 224, 224] in [0, 1]
        sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
        feats = torch.stack([vit_encode_image(sd_t, imgs[i], lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True)
                             for i in range(I)])
        _oracle_feats_cache[key] = (imgs, feats)
    return _oracle_feats_cache[key]


@pytest.mark.parametrize("text", ["random", "structured"])
@pytest.mark.parametrize("op", ["f16", "bf16"])
@pytest.mark.parametrize("I,V", [(64, 9), (24, 65)])
def test_end_to_end_agreement_with_fp32_oracle(jb, cuda_dev, lora_model, I, V, op, text):
    """The whole path against the fp32 oracle END TO END (oracle tower -> oracle MTA x3 -> oracle head), full
    12-layer LoRA tower, 64 images x 9 views and 24 images x 65 views (the headline N = 64).  A top-5 set can only differ
    where the oracle's 5th and 6th scores are closer than twice the logit deviation -- asserted per image.  The measured
    numbers are written to gpurun_out/e2e_agreement_<I>x<V>_<op>_<text>.json (copies under profiles/)."""
    import json
    import os
    from oracle import pipeline_image
    sd, lora, model = lora_model
    imgs, ref_feats = _oracle_feats(jb, sd, lora, I, V)
    if text == "structured":
        Ts = [torch.from_numpy(t) for t in jb.synth.make_structured_text_banks(ref_feats[:, 0].numpy(), seed=10)]
    else:
        Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_np = jb.synth.make_head(2, Ts[2].numpy())
    lp_t = tuple(torch.from_numpy(a) for a in lp_np)
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = lp_np
    ctx = jb.get_context(cuda_dev)
    prev = ctx.operand_type
    res = {}
    try:
        ctx.set_operand_type(op)
        for rank_by in ("cs5", "cs1"):
            hp = jb.HotPath(model, jb.TextBank(Ts[0], Ts[1], Ts[2], cuda_dev), lp, rank_by=rank_by)
            topk, feats, scores = hp.evaluate_base(torch.from_numpy(imgs).to(cuda_dev), return_feats=True, return_scores=True)
            topk, feats, scores = topk.cpu(), feats.cpu(), scores.cpu()
            agree, same_set, cos_min, dlogit, gap_at_flips = 0, 0, 1.0, 0.0, 0.0
            for i in range(I):
                f = ref_feats[i]
                cos_min = min(cos_min, float(_cos(feats[i], f).min()))
                t5, sc, _ = pipeline_image(f, f, Ts[0], Ts[1], Ts[2], lp_t, score=rank_by)
                ref_scores = sc[rank_by][0]
                d_i = float((scores[i] - ref_scores).abs().max())
                dlogit = max(dlogit, d_i)
                n = len(set(t5.tolist()) & set(topk[i].tolist()))
                agree += n
                same_set += n == 5
                if n < 5:
                    # a different top-5 set is only possible at a near-tie: the oracle's own 5th and 6th scores are closer
                    # than twice the logit deviation of this image
                    srt = ref_scores.sort(descending=True).values
                    gap = float(srt[4] - srt[5])
                    assert gap <= 2.0 * d_i + 1e-6, (i, gap, d_i)
                    assert n == 4, (i, n)                         # and then exactly one label is exchanged
                    gap_at_flips = max(gap_at_flips, gap)
            res[rank_by] = {"images": I, "views": V, "operands": op, "text": text,
                            "top5_label_agreement": agree / (5 * I), "identical_top5_sets": same_set / I,
                            "min_embedding_cosine": cos_min, "max_abs_logit_diff": dlogit,
                            "largest_oracle_5th_6th_gap_where_sets_differ": gap_at_flips}
    finally:
        ctx.set_operand_type(prev)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/e2e_agreement_{I}x{V}_{op}_{text}.json", "w") as fh:
        json.dump(res, fh, indent=1)
    print(res)
    b = E2E_BOUNDS[(op, text)]
    for rank_by, r in res.items():
        assert r["min_embedding_cosine"] >= 0.999, res
        assert b["dlogit"] is None or r["max_abs_logit_diff"] <= b["dlogit"], res
        assert r["top5_label_agreement"] >= b[rank_by], res
