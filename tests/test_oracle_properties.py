"""CPU: properties of the oracle itself (size-independent checks the GPU tests reuse at larger sizes)."""
import numpy as np
import pytest
import torch


def test_mta_mode_is_unit_and_view_order_invariant(jb):
    from oracle import solve_mta
    X = torch.from_numpy(jb.synth.make_unit_views(1, 1, 33))[0]
    T = torch.from_numpy(jb.synth.make_text_features(seed=2))
    m = solve_mta(X, T.t())
    assert m.shape == (1, 512) and abs(float(m.norm()) - 1) < 1e-6
    perm = torch.cat([torch.zeros(1, dtype=torch.long), 1 + torch.randperm(32, generator=torch.Generator().manual_seed(0))])
    m2 = solve_mta(X[perm], T.t())        # view 0 (the mode's start) stays; the crops are a set
    assert (m - m2).abs().max() < 1e-5


def test_mta_two_views_uses_k_of_one(jb):
    """N=1 crop (BASELINE config 5): the reference's k = int(0.3 * 1) = 0 averages nothing (NaN); the
    oracle defines k = max(k, 1)."""
    from oracle import solve_mta
    X = torch.from_numpy(jb.synth.make_unit_views(2, 1, 2))[0]
    T = torch.from_numpy(jb.synth.make_text_features(seed=2))
    m = solve_mta(X, T.t())
    assert torch.isfinite(m).all() and abs(float(m.norm()) - 1) < 1e-6


def test_mta_state_shapes(jb):
    from oracle import solve_mta
    X = torch.from_numpy(jb.synth.make_unit_views(3, 1, 17))[0]
    T = torch.from_numpy(jb.synth.make_text_features(seed=2))
    _, st = solve_mta(X, T.t(), return_state=True)
    assert st["affinity"].shape == (17, 17) and st["bandwidth"].shape == (17,)
    assert abs(float(st["y"].sum()) - 1) < 1e-5 and (st["bandwidth"] > 0).all()
    assert torch.allclose(st["affinity"], st["affinity"].t(), atol=1e-7)


def test_logit_normalize_semantics():
    from oracle import logit_normalize
    z = torch.randn(3, 403, generator=torch.Generator().manual_seed(0)) * 5 + 2
    out = logit_normalize(z)
    assert out.mean(dim=1).abs().max() < 1e-5                    # per-row mean removed
    assert torch.allclose(out, (z - z.mean(1, keepdim=True)) / z.std(unbiased=True))   # ONE global unbiased std
    one = logit_normalize(z[:1])
    assert abs(float(one.std(unbiased=True)) - 1) < 1e-5          # n = 1 (every reference call site): unit std


def test_topk_ties_lowest_index_first():
    from oracle import topk_lowest_index_first
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0]])
    assert topk_lowest_index_first(s, 5)[0].tolist() == [1, 2, 4, 3, 0]


def test_fusion_formulas(jb):
    from oracle import fuse_scores
    g = torch.Generator().manual_seed(1)
    m = [torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1) for _ in range(3)]
    T = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp = tuple(torch.from_numpy(a) for a in jb.synth.make_head(2, T[2].numpy()))
    sc = fuse_scores(m[0], m[1], m[2], T[0], T[1], T[2], lp)
    assert torch.allclose(sc["cs2"], (sc["cs"] + sc["cs1"]) / 2)
    assert torch.allclose(sc["cs4"], (sc["cs2"] + sc["cs3"]) / 2)
    assert torch.allclose(sc["cs5"], sc["cs4"] + 0.5 * sc["logits"])
    assert torch.allclose(sc["cs1"], 100.0 * m[0] @ T[0].t())


def test_vit_layernorm_matches_torch():
    from oracle.vit import layer_norm
    x = torch.randn(7, 768, generator=torch.Generator().manual_seed(0)) * 3 + 1
    w, b = torch.rand(768) + 0.5, torch.randn(768)
    assert torch.allclose(layer_norm(x, w, b), torch.nn.functional.layer_norm(x, (768,), w, b, 1e-5), atol=1e-5)


def test_folded_clip_normalisation_is_affine(jb):
    """tfm_clip fused on the device == (x - mean) / std applied first (reference test.py:1301, :1705)."""
    from oracle import vit_encode_image
    sd = jb.synth.make_vit_state_dict(seed=5, layers=1)
    img = jb.synth.make_views(6, 1, 2)[0]
    a = vit_encode_image(sd, img, apply_clip_norm=True)
    b = vit_encode_image(sd, jb.synth.clip_normalize(img))
    assert (a - b).abs().max() < 1e-5


def test_quantized_restatement_without_rounding_is_the_oracle(jb):
    """oracle/quantized.py (the tower with the GEMM operands rounded where the CUDA schedule rounds them) must be the
    fp32 oracle when nothing is rounded -- for the folded-LayerNorm schedule, the stand-alone one, and for LoRA in
    merged and in applied form (test.py:388-398: W x + b + s B (A x) is the same function as the merged weight)."""
    import torch
    from oracle import vit_encode_image
    from oracle.quantized import vit_encode_image_rounded
    sd = {k: torch.from_numpy(v) for k, v in jb.synth.make_vit_state_dict(seed=4, layers=2).items()}
    lora = jb.synth.make_lora(seed=5, layers=2, params=["q", "k", "v", "o"], b_std=0.3)
    imgs = jb.synth.make_views(3, 1, 3).reshape(3, 3, 224, 224)
    ref = vit_encode_image(sd, imgs, lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True)
    ref0 = vit_encode_image(sd, imgs, apply_clip_norm=True, normalize=True)
    assert (ref - ref0).abs().max() > 1e-3                      # the adapters matter
    for kw in ({"fold": True}, {"fold": False}, {"lora_mode": "applied"}):
        got = vit_encode_image_rounded(sd, imgs, lora=lora, scaling=0.5, act="f32", wgt="f32", **kw)
        assert (got - ref).abs().max() <= 2e-5, kw
    # and rounding the operands of the applied form to fp16 costs no more than it costs the merged form (x 2)
    e_m = (vit_encode_image_rounded(sd, imgs, lora=lora, scaling=0.5, act="f16", wgt="f16") - ref).abs().max()
    e_a = (vit_encode_image_rounded(sd, imgs, lora=lora, scaling=0.5, act="f16", wgt="f16", lora_mode="applied") - ref).abs().max()
    assert e_a <= 2 * e_m + 1e-5 and e_a <= 2e-3
