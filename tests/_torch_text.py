"""Plain-torch fp32 text tower over the product's parameter containers (CLIP.transformer etc.): a cross-check of the
HOST-side plumbing (parameter names, LoRA wrappers, dirty flags) that runs without a GPU.  Test code: it lived in
jclip/model.py in round 1 and was moved here because the product has no torch compute path.
Follows reference jclip/model.py:202-215 and test.py:388-398 (LoRA applied un-merged)."""
import math

import torch

from jclip_b200.runtime import as_torch


def _ln(x, ln, dev):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), ln.weight.torch(dev), ln.bias.torch(dev), 1e-5)


def _text_attention(attn, x, mask, dev):
    """Packed or LoRA-wrapped attention on [B,S,W] in torch (text tower only)."""
    W, H = attn.embed_dim, attn.num_heads
    w_in, b_in = attn.in_proj_weight.torch(dev), attn.in_proj_bias.torch(dev)
    qkv = []
    for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
        y = x @ w_in[j * W:(j + 1) * W].t() + b_in[j * W:(j + 1) * W]
        lin = getattr(attn, name, None)
        if lin is not None and getattr(lin, "lora_enabled", False):      # reference test.py:388-398
            y = y + (x @ (lin.w_lora_B.torch(dev) @ lin.w_lora_A.torch(dev)).t()) * lin.scaling
        qkv.append(y)
    B, S, _ = x.shape
    q, k, v = (t.view(B, S, H, W // H).transpose(1, 2) for t in qkv)
    a = (q @ k.transpose(-2, -1)) / math.sqrt(W // H) + mask[:S, :S]
    o = (torch.softmax(a, dim=-1) @ v).transpose(1, 2).reshape(B, S, W)
    y = o @ attn.out_proj.weight.torch(dev).t() + attn.out_proj.bias.torch(dev)
    lin = getattr(attn, "proj", None)
    if lin is not None and getattr(lin, "lora_enabled", False):
        y = y + (o @ (lin.w_lora_B.torch(dev) @ lin.w_lora_A.torch(dev)).t()) * lin.scaling
    return y


@torch.no_grad()
def encode_text_torch(model, text):
    # jclip/model.py:202-215
    self = model
    text = as_torch(text).long()
    dev = text.device
    x = self.token_embedding.weight.torch(dev)[text] + self.positional_embedding.torch(dev)
    mask = self.build_attention_mask().to(dev)
    for block in self.transformer.resblocks:
        x = x + _text_attention(block.attn, _ln(x, block.ln_1, dev), mask, dev)
        h = _ln(x, block.ln_2, dev) @ block.mlp.c_fc.weight.torch(dev).t() + block.mlp.c_fc.bias.torch(dev)
        h = h * torch.sigmoid(1.702 * h)
        x = x + h @ block.mlp.c_proj.weight.torch(dev).t() + block.mlp.c_proj.bias.torch(dev)
    x = _ln(x, self.ln_final, dev)
    eot = text.argmax(dim=-1)                                  # highest token id = EOT
    return x[torch.arange(x.shape[0], device=dev), eot] @ self.text_projection.torch(dev)
