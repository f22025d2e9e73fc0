"""Oracle: CLIP ViT image tower forward, fp32 on CPU.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py; parity unpinned).

Restates, line by line:
  * jclip/model.py:17-21   LayerNorm (Jittor nn.LayerNorm: eps 1e-5, affine, biased variance)
  * jclip/model.py:24-27   QuickGELU  x * sigmoid(1.702 x)
  * jclip/model.py:30-39   MLP c_fc -> gelu -> c_proj
  * jclip/model.py:42-62   ResidualAttentionBlock  x += attn(ln_1 x); x += mlp(ln_2 x)
  * jclip/model.py:104-126 VisionTransformer.execute
  * jclip/mha.py:55-83     scaled_dot_product_attention (no mask for vision, dropout 0)
  * jclip/mha.py:129-146   packed in-projection
  * test.py:277-398        LoRALayer / LinearLoRA (scaling alpha/sqrt(r); eval math Wx+b + s*x(BA)^T)
  * test.py:469-598        PlainMultiheadAttentionLoRA (packed in_proj split into q/k/v, rows 0:768, 768:1536, 1536:2304)
  * test.py:1301           tfm_clip = ImageNormalize(mean, std) applied to the batch before encode_image
"""
import math

import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # test.py:1301
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).cpu()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def layer_norm(x, weight, bias, eps=1e-5):
    # jclip/model.py:17-21 -> jittor nn.LayerNorm: biased variance over the last dim, eps inside sqrt
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def quick_gelu(x):
    # jclip/model.py:27
    return x * torch.sigmoid(1.702 * x)


def clip_normalize(images):
    # test.py:1301 T.ImageNormalize on a batched Var: (x - mean[c]) / std[c]
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (images - mean) / std


def lora_scaling(r, alpha):
    # test.py:288-289: alpha / sqrt(r)  (NOT alpha / r)
    return alpha / math.sqrt(r)


def merge_lora_into_state_dict(sd, lora, scaling, prefix="visual.transformer.resblocks."):
    """W' = W + s * B @ A per enabled projection (test.py:310-313 merge_BA, :388-398 eval math).

    `lora` maps vision layer index -> {'q_proj'|'k_proj'|'v_proj'|'proj': (A[r,in], B[out,r])}.
    Returns a new state dict (fp32 torch tensors); mathematically identical to the applied branch.
    """
    out = {k: _t(v).clone() for k, v in sd.items()}
    W = out[next(k for k in out if k.endswith("attn.in_proj_weight") and k.startswith(prefix))].shape[1]
    rows = {"q_proj": (0, W), "k_proj": (W, 2 * W), "v_proj": (2 * W, 3 * W)}   # test.py:491-501
    for layer, projs in (lora or {}).items():
        for name, (A, B) in projs.items():
            delta = scaling * (_t(B) @ _t(A))
            if name == "proj":
                out[f"{prefix}{layer}.attn.out_proj.weight"] += delta
            else:
                lo, hi = rows[name]
                out[f"{prefix}{layer}.attn.in_proj_weight"][lo:hi] += delta
    return out


def _attention(x, w_in, b_in, w_out, b_out, n_head, lora=None, scaling=0.0):
    """x: [B, S, W].  jclip/mha.py:201-466 hot branch / test.py:523-598."""
    B, S, W = x.shape
    d = W // n_head
    wq, wk, wv = w_in[:W], w_in[W:2 * W], w_in[2 * W:]
    bq, bk, bv = b_in[:W], b_in[W:2 * W], b_in[2 * W:]

    def lin(name, w, b):
        y = x @ w.t() + b                                   # nn.Linear
        if lora is not None and name in lora:               # test.py:388-398 (applied branch)
            A, Bm = lora[name]
            y = y + (x @ (_t(Bm) @ _t(A)).t()) * scaling
        return y

    q, k, v = lin("q_proj", wq, bq), lin("k_proj", wk, bk), lin("v_proj", wv, bv)
    # heads = contiguous 64-wide column blocks (jclip/mha.py:351-362, test.py:584-590)
    q = q.view(B, S, n_head, d).permute(0, 2, 1, 3)
    k = k.view(B, S, n_head, d).permute(0, 2, 1, 3)
    v = v.view(B, S, n_head, d).permute(0, 2, 1, 3)
    # jclip/mha.py:55-83: softmax(q k^T / sqrt(d)) v ; attn_mask None for vision (jclip/model.py:99)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(d))
    att = torch.softmax(att, dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(B, S, W)
    y = o @ w_out.t() + b_out                               # jclip/mha.py:461 / test.py:594
    if lora is not None and "proj" in lora:
        A, Bm = lora["proj"]
        y = y + (o @ (_t(Bm) @ _t(A)).t()) * scaling
    return y


@torch.no_grad()
def vit_encode_image(sd, images, lora=None, scaling=0.5, apply_clip_norm=False, normalize=False,
                     return_tokens=False):
    """CLIP.encode_image (jclip/model.py:199-200 -> :104-126), fp32.

    sd: state dict with the key names of jclip/model.py:235-285 (numpy or torch).
    images: [B,3,R,R] float32 (already CLIP-normalised unless apply_clip_norm).
    lora: {layer: {'q_proj': (A,B), ...}} applied un-merged as the reference's eval path does.
    normalize: also apply `f / f.norm(dim=-1, keepdim=True)` (test.py:1706).
    """
    x = _t(images)
    if apply_clip_norm:
        x = clip_normalize(x)
    g = lambda k: _t(sd[k])
    conv_w = g("visual.conv1.weight")
    width, _, P, _ = conv_w.shape
    n_head = width // 64                                     # jclip/model.py:152
    layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])

    x = torch.nn.functional.conv2d(x, conv_w, bias=None, stride=P)       # :105
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)           # :106-108
    cls = g("visual.class_embedding").view(1, 1, -1).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)                                       # :109-113
    x = x + g("visual.positional_embedding")                             # :114
    if "visual.VPT" in sd:
        # IVLP / VPT tower (jclip/model1.py:192-196): prompt tokens appended after pos-embed, before ln_pre;
        # the vision transformer has prompts_needed=0 (model1.py:175) so no layer replaces them
        vpt = g("visual.VPT")
        x = torch.cat([x, vpt.unsqueeze(0).expand(x.shape[0], -1, -1)], dim=1)
    x = layer_norm(x, g("visual.ln_pre.weight"), g("visual.ln_pre.bias"))  # :115
    for i in range(layers):                                              # :117-119 (layout permutes are no-ops here)
        p = f"visual.transformer.resblocks.{i}."
        h = layer_norm(x, g(p + "ln_1.weight"), g(p + "ln_1.bias"))
        x = x + _attention(h, g(p + "attn.in_proj_weight"), g(p + "attn.in_proj_bias"),
                           g(p + "attn.out_proj.weight"), g(p + "attn.out_proj.bias"), n_head,
                           lora=(lora or {}).get(i), scaling=scaling)    # :60
        h = layer_norm(x, g(p + "ln_2.weight"), g(p + "ln_2.bias"))
        h = quick_gelu(h @ g(p + "mlp.c_fc.weight").t() + g(p + "mlp.c_fc.bias"))
        x = x + (h @ g(p + "mlp.c_proj.weight").t() + g(p + "mlp.c_proj.bias"))   # :61
    if return_tokens:
        return x
    f = layer_norm(x[:, 0, :], g("visual.ln_post.weight"), g("visual.ln_post.bias"))  # :121
    f = f @ g("visual.proj")                                             # :123-124
    if normalize:
        f = f / f.norm(dim=-1, keepdim=True)                             # test.py:1706
    return f
