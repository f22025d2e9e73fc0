"""Host-side timing probe of HotPath.evaluate_image_stream (in stream order vs overlapped on a second stream): ms per step
and the host time of every TTAViews call.  python tools/probe_image_stream.py   (PIN=1: bind to the GPU's NUMA node first)"""
import os, sys, time, types
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import jclip_b200 as jb
dev = torch.device("cuda", 0)
if os.environ.get("PIN") == "1":
    print("numa cpus", jb.dist.bind_to_gpu_numa(0), flush=True)
sd = jb.synth.make_vit_state_dict(seed=0)
model = jb.jclip.build_model(sd)
texts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
lp = jb.Channel_LP()
lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = jb.synth.make_head(2, texts[2].numpy())
hp = jb.HotPath(model, jb.TextBank(*texts, dev), lp, rank_by="cs5")
rng = np.random.default_rng(0)
src = [rng.integers(0, 256, (375, 500, 3), dtype=np.uint8) for _ in range(128)]
gen = jb.TTAViews(n_crops=64, scale=(0.5, 1.0), seed=0, emit="patches")
orig_call = gen.__class__.__call__
times = []
def timed(self, *a, **k):
    t0 = time.perf_counter(); r = orig_call(self, *a, **k); times.append(time.perf_counter() - t0); return r
gen.__class__.__call__ = timed
for overlap in (False, True, False):
    for _ in hp.evaluate_image_stream((src for _ in range(2)), gen, overlap=overlap): pass
    torch.cuda.synchronize(); times.clear()
    t0 = time.perf_counter()
    for _ in hp.evaluate_image_stream((src for _ in range(8)), gen, overlap=overlap): pass
    torch.cuda.synchronize()
    print(f"overlap={overlap}: {(time.perf_counter()-t0)/8*1e3:.1f} ms/step; tta() host time per call {[round(t*1e3,1) for t in times]}", flush=True)
