#!/usr/bin/env python
"""Per-kernel CUDA-event profile (jcb_ctx_profile) of the reference's own call pattern: ONE image x 65 views per call,
graphs off so that every launch gets its event pair.  python tools/single_image_profile.py [views]"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import jclip_b200 as jb  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65
dev = torch.device("cuda", 0)
sd = jb.synth.make_vit_state_dict(seed=0)
model = jb.jclip.build_model(sd)
largs = types.SimpleNamespace(encoder="vision", position="all", params=["q", "k", "v"], r=4, alpha=1, dropout_rate=0.25,
                              backbone="ViT-B/32")
layers = jb.apply_lora(largs, model)
lora = jb.synth.make_lora(seed=7)
for i, layer in enumerate(layers):
    for name, (A, B) in lora[i].items():
        getattr(layer, name).w_lora_A.data = A
        getattr(layer, name).w_lora_B.data = B
texts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
lp = jb.Channel_LP()
lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = jb.synth.make_head(2, texts[2].numpy())
hp = jb.HotPath(model, jb.TextBank(*texts, dev), lp, rank_by="cs5")
ctx = jb.get_context(dev)
one = (jb.synth.make_views_torch(5, 1, V, dev) * 255).round_().to(torch.uint8)
ctx.set_graphs(False)
for _ in range(5):
    hp.evaluate_base(one, topk_to_host=True)
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N):
    hp.evaluate_base(one, topk_to_host=True)
wall = (time.perf_counter() - t0) / N * 1e3
ctx.profile_start()
for _ in range(N):
    hp.evaluate_base(one, topk_to_host=True)
prof = ctx.profile_stop()
tot = 0.0
print(f"1 image x {V} views: {wall:.3f} ms per blocking call (graphs off)")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    if v["launches"]:
        per_call = v["ms"] / N
        tot += per_call
        print(f"  {k:12s} {v['launches'] / N:5.1f} launches/call  {1e3 * per_call:8.1f} us/call  {1e3 * v['ms'] / max(v['timed_launches'], 1):7.1f} us/launch")
print(f"  sum of kernel times {1e3 * tot:.1f} us of {1e3 * wall:.1f} us wall")
