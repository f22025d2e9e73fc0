"""Where does the end-to-end logit deviation of the 16-bit tensor-core path come from?  TEST INFRASTRUCTURE ONLY.

    python -m oracle.attribution [--images 64 --views 9] [--out profiles/r02_error_attribution.json]

Runs the fp32 oracle pipeline (tower -> solve_mta x3 -> head) and the same pipeline with the tower's GEMM operands
rounded as the CUDA schedule rounds them (oracle/quantized.py), on the inputs of
tests/test_gpu_fullsize.py::test_end_to_end_agreement_with_fp32_oracle, for each combination of activation / weight
operand type.  CPU only; no GPU library involved.  Reports min embedding cosine, max |dlogit| for cs1 / cs5 and the
top-5 label agreement with the fp32 oracle (north star: 1e-2, 99.5 %).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--views", type=int, default=9)
    ap.add_argument("--out", default="")
    ap.add_argument("--variants", default="bf16/bf16,f16/f16,f32/bf16,bf16/f32,f32/f16,f16/f32,bf16/f16,f16/bf16")
    ap.add_argument("--text", default="random", choices=["random", "structured"])
    a = ap.parse_args()
    import jclip_b200 as jb
    from oracle import pipeline_image, vit_encode_image
    from oracle.quantized import vit_encode_image_rounded

    torch.set_num_threads(os.cpu_count() or 1)
    I, V = a.images, a.views
    sd = {k: torch.from_numpy(v) for k, v in jb.synth.make_vit_state_dict(seed=0).items()}
    lora = jb.synth.make_lora(seed=7, b_std=0.05)
    imgs = jb.synth.make_views(21, I, V)
    def towers(encode):
        return torch.stack([encode(imgs[i]) for i in range(I)])          # [I, V, 512] unit view embeddings

    t0 = time.time()
    ref_feats = towers(lambda x: vit_encode_image(sd, x, lora=lora, scaling=0.5, apply_clip_norm=True, normalize=True))
    print(f"fp32 oracle tower: {time.time() - t0:.1f} s", flush=True)
    if a.text == "structured":
        Ts = [torch.from_numpy(t) for t in jb.synth.make_structured_text_banks(ref_feats[:, 0].numpy(), seed=10)]
    else:
        Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_t = tuple(torch.from_numpy(x) for x in jb.synth.make_head(2, Ts[2].numpy()))

    def heads(feats):
        scores, tops = {"cs1": [], "cs5": []}, {"cs1": [], "cs5": []}
        for i in range(I):
            for s in ("cs1", "cs5"):
                t5, sc, _ = pipeline_image(feats[i], feats[i], Ts[0], Ts[1], Ts[2], lp_t, score=s)
                scores[s].append(sc[s][0])
                tops[s].append(set(t5.tolist()))
        return feats, {s: torch.stack(v) for s, v in scores.items()}, tops

    ref = heads(ref_feats)
    gaps = {s: float((ref[1][s].sort(descending=True).values[:, 4] - ref[1][s].sort(descending=True).values[:, 5]).min())
            for s in ("cs1", "cs5")}
    print(f"smallest 5th/6th score gap of the fp32 oracle: {gaps}", flush=True)
    run = lambda encode: heads(towers(encode))
    out = {"images": I, "views": V, "text": a.text, "oracle_min_gap_5th_6th": gaps, "variants": {}}
    for var in a.variants.split(","):
        # "act/wgt" or "act/wgt/applied" (LoRA as low-rank GEMMs, stand-alone LayerNorm schedule: jcb_ctx_set_lora_mode)
        act, wgt, *mode = var.split("/")
        lora_mode = mode[0] if mode else "merged"
        t0 = time.time()
        got = run(lambda x: vit_encode_image_rounded(sd, x, lora=lora, scaling=0.5, act=act, wgt=wgt, lora_mode=lora_mode))
        cos = torch.nn.functional.cosine_similarity(got[0].double(), ref[0].double(), dim=-1)
        row = {"min_embedding_cosine": float(cos.min()), "max_embedding_l2": float((got[0] - ref[0]).norm(dim=-1).max())}
        for s in ("cs1", "cs5"):
            d = (got[1][s] - ref[1][s]).abs()
            row[s] = {"max_abs_logit_diff": float(d.max()), "mean_abs_logit_diff": float(d.mean()),
                      "top5_label_agreement": sum(len(x & y) for x, y in zip(got[2][s], ref[2][s])) / (5 * I)}
        out["variants"][f"act={act},wgt={wgt}" + (f",lora={lora_mode}" if mode else "")] = row
        print(f"act={act} wgt={wgt} lora={lora_mode} ({time.time() - t0:.1f} s): {json.dumps(row)}", flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
