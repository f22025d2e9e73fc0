"""GPU parity of `encode_text` (CLIP text tower: 77 tokens, width 512, 8 heads, causal mask, EOT gather,
text_projection; LoRA on q/k/v of the text blocks) against the fp32 CPU oracle (oracle/text.py), through the
reference-facing API (`model.encode_text`).  Tolerance: cosine >= 0.9995 (bf16 GEMM operands, fp32 elsewhere)."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double(), b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))


def test_encode_text_matches_oracle(jb, cuda_dev):
    from oracle import text_encode
    sd = jb.synth.make_vit_state_dict(seed=5, layers=1, text_layers=12)
    model = jb.jclip.build_model(sd)
    tok = jb.synth.make_tokens(6, 37, vocab=64, max_len=74)
    tok[0, 1:76] = 3; tok[0, 76] = 63                         # a sequence that fills the whole context
    ref = text_encode(sd, tok)
    out = model.encode_text(torch.from_numpy(tok).to(cuda_dev)).cpu()
    assert out.shape == (37, 512)
    assert _cos(out, ref).min() >= 0.9995, _cos(out, ref).min()
    out_h = model.encode_text(tok)                            # host tokens are uploaded
    assert torch.equal(out_h.cpu(), out)
    from _torch_text import encode_text_torch
    assert (_cos(encode_text_torch(model, torch.from_numpy(tok)), ref) > 0.99999).all()


def test_encode_text_lora_and_refresh(jb, cuda_dev):
    from oracle import text_encode
    sd = jb.synth.make_vit_state_dict(seed=6, layers=1, text_layers=3)
    model = jb.jclip.build_model(sd)
    args = types.SimpleNamespace(encoder="text", position="all", params=["q", "k", "v"], r=4, alpha=1,
                                 dropout_rate=0.25, backbone="ViT-B/32")
    layers = jb.apply_lora(args, model)
    assert len(layers) == 3
    tok = torch.from_numpy(jb.synth.make_tokens(7, 9, vocab=64)).to(cuda_dev)
    f0 = model.encode_text(tok, normalize=True).clone()
    lora = jb.synth.make_lora(seed=9, layers=3, width=512, b_std=0.3)
    for i, layer in enumerate(layers):
        for name, (A, B) in lora[i].items():
            getattr(layer, name).w_lora_A.data = A
            getattr(layer, name).w_lora_B.data = B
    f1 = model.encode_text(tok, normalize=True).cpu()
    ref = text_encode(sd, tok.cpu().numpy(), lora=lora, scaling=0.5, normalize=True)
    assert _cos(f1, ref).min() >= 0.9995
    assert _cos(f1, f0.cpu()).min() < 0.999                   # the adapters matter
    assert (f1.norm(dim=-1) - 1).abs().max() < 1e-5


def test_causal_attention_kernel(jb, cuda_dev):
    """The text tower's attention: T = 77 padded to 80, key j visible to query i only if j <= i."""
    import math
    from ctypes import c_void_p
    # there is no stand-alone C-ABI entry for the causal variant; the tower with 1 block is the probe: compare the
    # full-context sequence against a torch reference at the embedding level instead (covered above), and check here
    # that a change in a LATER token never alters the EOT embedding of an earlier EOT position
    sd = jb.synth.make_vit_state_dict(seed=8, layers=1, text_layers=2)
    model = jb.jclip.build_model(sd)
    tok = jb.synth.make_tokens(9, 4, vocab=64, max_len=10)
    a = model.encode_text(torch.from_numpy(tok).to(cuda_dev)).cpu()
    tok2 = tok.copy()
    tok2[:, 40:60] = 5                                        # garbage after the EOT (ids below the EOT id)
    b = model.encode_text(torch.from_numpy(tok2).to(cuda_dev)).cpu()
    assert torch.equal(a, b)


def test_clip_classifier(jb, cuda_dev):
    """reference test.py:920-940 on pre-tokenised templates: per class mean of unit embeddings, re-normalised."""
    from oracle import text_encode
    sd = jb.synth.make_vit_state_dict(seed=8, layers=1, text_layers=2)
    model = jb.jclip.build_model(sd)
    templates = {c: jb.synth.make_tokens(100 + c, 2 + c % 3, vocab=64) for c in range(7)}
    W = jb.clip_classifier(templates, model)
    assert W.shape == (1, 7, 512) and W.is_cuda
    for c, tok in templates.items():
        e = text_encode(sd, tok, normalize=True).mean(dim=0)
        e = e / e.norm()
        assert _cos(W[0, c].cpu(), e) >= 0.9995
    assert (W[0].norm(dim=-1).cpu() - 1).abs().max() < 1e-5
