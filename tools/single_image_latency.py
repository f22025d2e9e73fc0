#!/usr/bin/env python
"""Latency of the reference's own call pattern -- ONE image x (N + 1) views per blocking call (test.py:1692-1742) --
with and without CUDA graphs (jcb_ctx_set_graphs), for N in {1, 16, 64}.  python tools/single_image_latency.py"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import jclip_b200 as jb  # noqa: E402

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
for n_crops in (1, 16, 64):
    V = n_crops + 1
    one = (jb.synth.make_views_torch(5, 1, V, dev) * 255).round_().to(torch.uint8)
    row = {}
    for graphs in (False, True):
        ctx.set_graphs(graphs)
        for _ in range(5):
            hp.evaluate_base(one, topk_to_host=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(100):
            hp.evaluate_base(one, topk_to_host=True)
        row[graphs] = (time.perf_counter() - t0) / 100 * 1e3
    print(f"1 image x {V} views: {row[False]:.3f} ms per call without graphs, {row[True]:.3f} ms with ({ctx.graph_stats()})")
