"""GPU: the three LayerNorm schedules of the tower (JCB_LN_FOLD = 0 stand-alone LayerNorm passes, 1 = ln_1 folded
into c_proj's epilogue + the QKV GEMM, 2 = ln_2 folded as well (default)) all meet the embedding tolerance against
the fp32 oracle -- for both operand types, on random-init weights AND on weights that give the residual stream the
statistics of trained CLIP towers (synth.make_vit_state_dict(trained_like=...): row mean 20 x the spread, outlier
channels 100 x the rest), where a fold that rounds the RAW residual copy loses the tolerance (oracle/quantized.py:
cosine 0.99976 for bf16) and the centred copy does not.  The mode is read when the context is created, hence one
subprocess per mode.  Reference: jclip/model.py:17-21, :59-62."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import sys, types
sys.path.insert(0, %r)
import numpy as np, torch
import jclip_b200 as jb
from oracle import vit_encode_image, text_encode
dev = torch.device("cuda", 0)
ctx = jb.get_context(dev)
worst = {}
for trained_like in (False, "offset", "outliers", "both"):
    sd = jb.synth.make_vit_state_dict(seed=1, text_layers=2, trained_like=trained_like)
    model = jb.jclip.build_model(sd)
    args = types.SimpleNamespace(encoder="both", position="all", params=["q", "k", "v"], r=4, alpha=1, dropout_rate=0.25,
                                 backbone="ViT-B/32")
    layers = jb.apply_lora(args, model)
    lora_t = jb.synth.make_lora(seed=5, layers=2, width=512, b_std=0.3)
    lora_v = jb.synth.make_lora(seed=7, b_std=0.3)
    for i, layer in enumerate(layers):
        src = lora_t[i] if i < 2 else lora_v[i - 2]
        for name, (A, B) in src.items():
            getattr(layer, name).w_lora_A.data = A
            getattr(layer, name).w_lora_B.data = B
    imgs = jb.synth.make_views(5, 1, 6).reshape(6, 3, 224, 224)
    ref = vit_encode_image(sd, imgs, lora=lora_v, scaling=0.5, apply_clip_norm=True, normalize=True)
    tok = jb.synth.make_tokens(3, 11, vocab=64)
    reft = text_encode(sd, tok, lora=lora_t, scaling=0.5, normalize=True)
    for op in ("bf16", "f16"):
        ctx.set_operand_type(op)
        out = model.visual(torch.from_numpy(imgs).to(dev), apply_clip_norm=True, normalize=True).cpu()
        cos = torch.nn.functional.cosine_similarity(out.double(), ref.double(), dim=-1).min().item()
        l2 = (out - ref).norm(dim=-1).max().item()
        outt = model.encode_text(torch.from_numpy(tok).to(dev), normalize=True).cpu()
        cost = torch.nn.functional.cosine_similarity(outt.double(), reft.double(), dim=-1).min().item()
        print("RESULT", trained_like, op, cos, l2, cost)
        # north star: cosine >= 0.999; asserted tighter.  fp16 operands: 8x less rounding -> 1 - cos 64x smaller
        assert cos >= (0.9995 if op == "bf16" else 0.99999) and cost >= (0.9995 if op == "bf16" else 0.99999), (trained_like, op, cos, cost)
''' % ROOT


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_layernorm_schedules(mode):
    env = dict(os.environ, JCB_LN_FOLD=mode)
    r = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RESULT") == 8
