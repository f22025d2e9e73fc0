"""LoRA in APPLIED form (`jcb_ctx_set_lora_mode(JCB_LORA_APPLIED)`, `Context.set_lora_mode("applied")`):
y = W x + b + s B (A x), the branch the reference evaluates in eval mode (test.py:388-398; its Jittor `eval()` never
merges), as low-rank tcgen05 GEMMs: U = x [A_q; A_k; A_v]^T, then the projection GEMM accumulates
U [s B_q | s B_k | s B_v]^T into the same TMEM tile through its second TMA operand pair.

Checked (i) at the kernel level: C = A B^T + A2 B2^T of `jcb_gemm` against torch fp32 on the same 16-bit operands, for
the epilogues the applied schedule uses, both operand types; (ii) through `encode_image` / `encode_text` against the
fp32 oracle, which evaluates the un-merged form as well: the same cosine bounds as the merged mode (>= 0.9995 bf16,
>= 0.99999 fp16; north star >= 0.999), and applied vs merged embeddings within 2e-3; (iii) the error behaviour.
"""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(params=("q", "k", "v"), encoder="vision", r=4):
    return types.SimpleNamespace(encoder=encoder, position="all", params=list(params), r=r, alpha=1,
                                 dropout_rate=0.25, backbone="ViT-B/32")


def _cos(a, b):
    a, b = a.double(), b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))


@pytest.fixture
def applied(jb, cuda_dev):
    """The context in applied mode for the duration of one test, merged (the default) afterwards."""
    ctx = jb.get_context(cuda_dev)
    assert ctx.lora_mode == "merged"
    ctx.set_lora_mode("applied")
    yield ctx
    ctx.set_lora_mode("merged")


def _set_adapters(layers, lora):
    for i, layer in enumerate(layers):
        for name, (A, B) in lora[i].items():
            getattr(layer, name).w_lora_A.data = A
            getattr(layer, name).w_lora_B.data = B


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("epi,N", [(0, 2304), (2, 768), (4, 768), (0, 128)])
def test_gemm_second_operand_pair(jb, cuda_dev, dtype, epi, N):
    """C = A B^T + A2[:, :64] B2^T (+ bias; += for the residual epilogue): A2 is the first 64 columns of a
    128-column matrix, exactly how the tower passes the low-rank intermediate."""
    g = torch.Generator(device="cpu").manual_seed(11 + epi + N)
    M, K, K2 = 777, 768, 64
    A = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * 0.05).to(dtype).to(cuda_dev)
    U = (torch.randn(M, 128, generator=g) * 0.5).to(dtype).to(cuda_dev)
    B2 = (torch.randn(N, K2, generator=g) * 0.2).to(dtype).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ref = A.float() @ B.float().T + U[:, :K2].float() @ B2.float().T + bias
    base = A.float() @ B.float().T + bias
    assert (ref - base).abs().max() > 0.5                    # the second pair matters
    if epi == 0:
        out = torch.empty(M, N, dtype=dtype, device=cuda_dev)
    elif epi == 2:
        res = torch.randn(M, N, generator=g).to(cuda_dev)
        out = res.clone()
        ref = ref + res
    else:
        out = torch.full((M, N), float("nan"), device=cuda_dev)
    jb.blocks.gemm(A, B, out, epi, bias=bias, A2=U, B2=B2, K2=K2)
    ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    tol = 2e-3 * np.sqrt((K + K2) / 768) + 1e-3 * ref.abs() + (2 * ulp * ref.abs() if epi == 0 else 0)
    assert ((out.float() - ref).abs() <= tol).all(), (out.float() - ref).abs().max()
    # and without the second pair the same call gives the base product (K2 = 0 path of the same kernel)
    out0 = torch.zeros_like(out) if epi != 2 else res.clone()
    jb.blocks.gemm(A, B, out0, epi, bias=bias)
    ref0 = base + (res if epi == 2 else 0)
    assert ((out0.float() - ref0).abs() <= 2e-3 + 1e-3 * ref0.abs() + (2 * ulp * ref0.abs() if epi == 0 else 0)).all()


def test_gemm_second_operand_pair_rejects_bad_k2(jb, cuda_dev):
    A = torch.zeros(128, 768, dtype=torch.float16, device=cuda_dev)
    B = torch.zeros(128, 768, dtype=torch.float16, device=cuda_dev)
    U = torch.zeros(128, 128, dtype=torch.float16, device=cuda_dev)
    B2 = torch.zeros(128, 64, dtype=torch.float16, device=cuda_dev)
    out = torch.zeros(128, 128, dtype=torch.float16, device=cuda_dev)
    with pytest.raises(RuntimeError):
        jb.blocks.gemm(A, B, out, 0, A2=U, B2=B2, K2=48)      # not a multiple of the 64-wide k-block
    with pytest.raises(RuntimeError):
        jb.blocks.gemm(A, B, out, 0, A2=U, B2=B2, K2=128)     # wider than B2's rows


@pytest.mark.parametrize("op", ["f16", "bf16"])
@pytest.mark.parametrize("params", [("q", "k", "v"), ("q", "v"), ("q", "k", "v", "o"), ("o",)])
def test_encode_image_lora_applied_matches_oracle(jb, cuda_dev, applied, params, op):
    from oracle import vit_encode_image
    ctx = applied
    prev = ctx.operand_type
    ctx.set_operand_type(op)
    try:
        sd = jb.synth.make_vit_state_dict(seed=1)
        model = jb.jclip.build_model(sd)
        layers = jb.apply_lora(_args(params), model)
        lora = jb.synth.make_lora(seed=7, params=list(params), b_std=0.3)
        _set_adapters(layers, lora)
        imgs = jb.synth.make_views(5, 1, 6).reshape(6, 3, 224, 224)
        ref = vit_encode_image(sd, imgs, lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True)
        ref0 = vit_encode_image(sd, imgs, lora=None, apply_clip_norm=True, normalize=True)
        x = torch.from_numpy(imgs).to(cuda_dev)
        out = model.visual(x, apply_clip_norm=True, normalize=True).cpu()
        assert ctx.lib.jcb_vit_lora_mode(model.visual._vit) == jb._capi.LORA_APPLIED
        cos = _cos(out, ref)
        assert cos.min() >= (0.99999 if op == "f16" else 0.9995), cos
        assert (out.norm(dim=-1) - 1).abs().max() < 1e-5
        assert _cos(ref, ref0).min() < 0.9999                 # the adapters matter
        assert (_cos(out, ref) > _cos(out, ref0)).all()
        # merged mode on the same model: re-packed lazily, same embeddings to operand rounding
        ctx.set_lora_mode("merged")
        out_m = model.visual(x, apply_clip_norm=True, normalize=True).cpu()
        assert ctx.lib.jcb_vit_lora_mode(model.visual._vit) == jb._capi.LORA_MERGED
        assert (out - out_m).abs().max() <= (2e-3 if op == "f16" else 1e-2)
        ctx.set_lora_mode("applied")
        out_a = model.visual(x, apply_clip_norm=True, normalize=True).cpu()
        assert torch.equal(out_a, out)                        # back and forth: bit-identical
    finally:
        ctx.set_operand_type(prev)


def test_applied_mode_without_adapters_is_the_plain_tower(jb, cuda_dev, applied):
    """No adapters: the applied schedule is the stand-alone-LayerNorm schedule of the zero-shot tower."""
    from oracle import vit_encode_image
    sd = jb.synth.make_vit_state_dict(seed=0)
    model = jb.jclip.build_model(sd)
    imgs = jb.synth.clip_normalize(jb.synth.make_views(3, 2, 4).reshape(8, 3, 224, 224))
    ref = vit_encode_image(sd, imgs)
    out = model.encode_image(torch.from_numpy(imgs).to(cuda_dev)).cpu()
    assert _cos(out, ref).min() >= 0.99999


def test_adapter_update_in_applied_mode(jb, cuda_dev, applied):
    """New adapter values reach the device on the next call (the reference mutates w_lora_A / w_lora_B in place,
    test.py:723-733)."""
    from oracle import vit_encode_image
    sd = jb.synth.make_vit_state_dict(seed=2, layers=2)
    model = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(("q", "k", "v")), model)
    imgs = jb.synth.make_views(6, 1, 4).reshape(4, 3, 224, 224)
    x = torch.from_numpy(imgs).to(cuda_dev)
    outs = []
    for seed in (3, 4):
        lora = jb.synth.make_lora(seed=seed, layers=2, b_std=0.3)
        _set_adapters(layers, lora)
        out = model.visual(x, apply_clip_norm=True, normalize=True).cpu()
        ref = vit_encode_image(sd, imgs, lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True)
        assert _cos(out, ref).min() >= 0.99999
        outs.append(out)
    assert (outs[0] - outs[1]).abs().max() > 1e-3


def test_rank_sum_above_64_is_rejected(jb, cuda_dev, applied):
    sd = jb.synth.make_vit_state_dict(seed=2, layers=1)
    model = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(("q", "k", "v"), r=32), model)            # 96 > 64
    lora = jb.synth.make_lora(seed=3, layers=1, r=32, b_std=0.3)
    _set_adapters(layers, lora)
    x = torch.zeros(1, 3, 224, 224, device=cuda_dev)
    with pytest.raises(RuntimeError, match="ranks"):
        model.visual(x)
    applied.set_lora_mode("merged")                                         # merged takes any rank
    assert model.visual(x).shape == (1, 512)


def test_rank_16_qkv_fills_48_of_64_columns(jb, cuda_dev, applied):
    from oracle import vit_encode_image
    sd = jb.synth.make_vit_state_dict(seed=2, layers=2)
    model = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(("q", "k", "v", "o"), r=16), model)
    lora = jb.synth.make_lora(seed=5, layers=2, r=16, params=["q", "k", "v", "o"], b_std=0.1)
    _set_adapters(layers, lora)
    imgs = jb.synth.make_views(6, 1, 3).reshape(3, 3, 224, 224)
    ref = vit_encode_image(sd, imgs, lora=lora, scaling=1.0 / 4.0, apply_clip_norm=True, normalize=True)
    out = model.visual(torch.from_numpy(imgs).to(cuda_dev), apply_clip_norm=True, normalize=True).cpu()
    assert _cos(out, ref).min() >= 0.99999


def test_encode_text_lora_applied(jb, cuda_dev, applied):
    from oracle import text_encode
    sd = jb.synth.make_vit_state_dict(seed=6, layers=1, text_layers=3)
    model = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(("q", "k", "v"), encoder="text"), model)
    lora = jb.synth.make_lora(seed=9, layers=3, width=512, b_std=0.3)
    _set_adapters(layers, lora)
    tok = torch.from_numpy(jb.synth.make_tokens(7, 9, vocab=64)).to(cuda_dev)
    f1 = model.encode_text(tok, normalize=True).cpu()
    ref = text_encode(sd, tok.cpu().numpy(), lora=lora, scaling=0.5, normalize=True)
    ref0 = text_encode(sd, tok.cpu().numpy(), normalize=True)
    assert applied.lib.jcb_text_lora_mode(model._text) == jb._capi.LORA_APPLIED
    assert _cos(f1, ref).min() >= 0.99999
    assert _cos(ref, ref0).min() < 0.999


def test_pipeline_in_applied_mode_matches_merged(jb, cuda_dev, applied):
    """The whole hot path (tower -> MTA x3 -> head -> top-5) with applied adapters: same labels as merged."""
    sd = jb.synth.make_vit_state_dict(seed=0)
    model = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(("q", "k", "v")), model)
    _set_adapters(layers, jb.synth.make_lora(seed=7, b_std=0.02))
    I, V = 4, 9
    imgs = torch.from_numpy(jb.synth.make_views(1, I, V)).to(cuda_dev)
    texts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_np = jb.synth.make_head(2, texts[2].numpy())
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = lp_np
    hp = jb.HotPath(model, jb.TextBank(*texts, cuda_dev), lp, rank_by="cs1")
    top_a, sc_a = hp.evaluate_base(imgs, return_scores=True)
    top_a, sc_a = top_a.cpu(), sc_a.cpu()
    applied.set_lora_mode("merged")
    top_m, sc_m = hp.evaluate_base(imgs, return_scores=True)
    assert (sc_a - sc_m.cpu()).abs().max() <= 1e-2
    assert torch.equal(top_a, top_m.cpu())
