#!/usr/bin/env python
"""MTA inside the pipeline call (three banks on one feature tensor, as in the headline step) with a one-layer tower:
    python tools/mta_probe.py [images] [views]      -> ms per call of the MTA kernels (library profiling hook)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import jclip_b200 as jb  # noqa: E402

I = int(sys.argv[1]) if len(sys.argv) > 1 else 128
V = int(sys.argv[2]) if len(sys.argv) > 2 else 65
dev = torch.device("cuda", 0)
model = jb.jclip.build_model(jb.synth.make_vit_state_dict(seed=4, layers=1))
Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
lp = jb.Channel_LP()
lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = jb.synth.make_head(2, Ts[2].numpy())
hp = jb.HotPath(model, jb.TextBank(Ts[0], Ts[1], Ts[2], dev), lp, rank_by="cs5")
imgs = (jb.synth.make_views_torch(51, I, V, dev) * 255).round_().to(torch.uint8)
ctx = jb.get_context(dev)
ctx.set_graphs(False)
for _ in range(3):
    hp.evaluate_base(imgs)
ctx.profile_start()
for _ in range(5):
    hp.evaluate_base(imgs)
prof = ctx.profile_stop()
print(f"I={I} V={V}: " + "  ".join(f"{k} {v['ms'] / 5:.3f} ms" for k, v in prof.items() if k in ("mta", "head", "tail")))
