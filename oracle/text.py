"""Oracle: CLIP text tower forward, fp32 on CPU.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates jclip/model.py:202-215 (CLIP.encode_text: token + positional embedding, causal transformer,
ln_final, EOT gather by `text.argmax(dim=-1)`, text_projection), :189-193 (build_attention_mask) and the
LoRA eval math of test.py:388-398 on the text blocks (apply_lora with encoder 'text' / 'both',
test.py:611-623).  Pinned against the reference source run on the Jittor stand-in (tests/golden).
"""
import math

import torch

from .vit import _t, layer_norm, quick_gelu


@torch.no_grad()
def text_encode(sd, tokens, lora=None, scaling=0.5, normalize=False):
    """sd: state dict (reference key names); tokens [n, context] int64; lora {layer: {'q_proj': (A, B), ...}}."""
    g = lambda k: _t(sd[k])
    tok = torch.as_tensor(tokens).long()
    x = g("token_embedding.weight")[tok] + g("positional_embedding")           # :203-205
    n, S, W = x.shape
    H = W // 64                                                                 # jclip/model.py:268
    layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
    mask = torch.triu(torch.full((S, S), float("-inf")), 1)                    # :189-193
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        h = layer_norm(x, g(p + "ln_1.weight"), g(p + "ln_1.bias"))
        w_in, b_in = g(p + "attn.in_proj_weight"), g(p + "attn.in_proj_bias")
        ad = (lora or {}).get(i, {})
        qkv = []
        for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
            y = h @ w_in[j * W:(j + 1) * W].t() + b_in[j * W:(j + 1) * W]
            if name in ad:
                A, B = ad[name]
                y = y + (h @ (_t(B) @ _t(A)).t()) * scaling
            qkv.append(y.view(n, S, H, 64).permute(0, 2, 1, 3))
        att = torch.softmax(qkv[0] @ qkv[1].transpose(-2, -1) / math.sqrt(64) + mask, dim=-1)
        o = (att @ qkv[2]).permute(0, 2, 1, 3).reshape(n, S, W)
        y = o @ g(p + "attn.out_proj.weight").t() + g(p + "attn.out_proj.bias")
        if "proj" in ad:
            A, B = ad["proj"]
            y = y + (o @ (_t(B) @ _t(A)).t()) * scaling
        x = x + y
        h = layer_norm(x, g(p + "ln_2.weight"), g(p + "ln_2.bias"))
        h = quick_gelu(h @ g(p + "mlp.c_fc.weight").t() + g(p + "mlp.c_fc.bias"))
        x = x + h @ g(p + "mlp.c_proj.weight").t() + g(p + "mlp.c_proj.bias")
    x = layer_norm(x, g("ln_final.weight"), g("ln_final.bias"))               # :211
    f = x[torch.arange(n), tok.argmax(dim=-1)] @ g("text_projection")          # :213-214 (EOT = highest id)
    if normalize:
        f = f / f.norm(dim=-1, keepdim=True)
    return f
