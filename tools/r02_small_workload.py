#!/usr/bin/env python
"""Workload for the round-2 ncu capture of the kernels that are new this round: the text tower's tcgen05 attention
(77 tokens, causal, 8 heads; 806 sequences = 403 classes x 2 templates), the view generator writing the conv1 patch matrix
(16 images x 65 views), and the MTA kernels of one bench step's worth of embeddings (128 images x 65 views x 3 banks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jclip_b200 as jb  # noqa: E402

dev = torch.device("cuda", 0)
sd = jb.synth.make_vit_state_dict(seed=5, layers=1, text_layers=12)
model = jb.jclip.build_model(sd)
tok = torch.from_numpy(jb.synth.make_tokens(6, 806, vocab=64, max_len=74)).to(dev)
for _ in range(2):
    out = model.encode_text(tok, normalize=True)
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, (375, 500, 3), dtype=np.uint8) for _ in range(16)]
gen = jb.TTAViews(n_crops=64, seed=0, emit="patches")
for _ in range(2):
    p = gen(imgs)
feats = torch.from_numpy(jb.synth.make_unit_views(0, 128, 65)).to(dev)
Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)).t().contiguous().to(dev) for i in range(3)]
for _ in range(2):
    for T in Ts:
        m = jb.solve_mta_batched(feats, T)
torch.cuda.synchronize()
print("ok", out.shape, p.shape, m.shape)
