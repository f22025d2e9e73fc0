#!/usr/bin/env python
"""Generates tests/golden/ref_on_shim.npz by EXECUTING THE REFERENCE'S OWN SOURCE for the hot path on
the torch-backed Jittor stand-in (oracle/jt_shim.py).  TEST INFRASTRUCTURE ONLY.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py [--reference /root/reference] [--out tests/golden/ref_on_shim.npz]

What is executed, unmodified, from the reference checkout:
  * jclip/mha.py, jclip/model.py and jclip/model1.py (imported as modules): build_model -> CLIP.encode_image,
    for the plain tower and for the 54-token IVLP / VPT tower
  * from test.py, by source extraction (the file cannot be imported: it reads `text_template/` and
    dataset files at import time): LoRALayer, LinearLoRA, _canonical_mask,
    scaled_dot_product_attention, _none_or_dtype, PlainMultiheadAttentionLoRA, apply_lora,
    INDEX_POSITIONS_TEXT / _VISION, Channel_LP, logit_normalize, gaussian_kernel, cdist, solve_mta
  * from ood.py: its solve_mta twin (returns 100 * mode @ text)
No reference source is copied into this repository; only the numerical outputs are stored.

Inputs come from the seeded generators in jittor-clip-fewshot_b200/synth.py; the fixture stores the outputs plus a
checksum of every input so a drift of the generators is detected by tests/test_oracle_golden.py.
"""
import argparse
import ast
import importlib
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import jt_shim  # noqa: E402


def extract(path, names):
    """Source text of the named top-level defs / classes / assignments of a Python file."""
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    found = {}
    for node in tree.body:
        nm = None
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            nm = node.name
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            nm = node.targets[0].id
        if nm in names:
            found[nm] = ast.get_source_segment(src, node)      # last definition wins, as at import time
    missing = [n for n in names if n not in found]
    if missing:
        raise RuntimeError(f"{path}: not found: {missing}")
    return "\n\n".join(found[n] for n in names)


def load_ref_jclip(ref):
    """Import reference jclip/mha.py + jclip/model.py as package `_refjclip` WITHOUT running
    jclip/__init__.py (which pulls the tokenizer, ftfy and jittor.transform)."""
    pkg = types.ModuleType("_refjclip")
    pkg.__path__ = [os.path.join(ref, "jclip")]
    sys.modules["_refjclip"] = pkg
    for name in ("mha", "model", "model1"):
        spec = importlib.util.spec_from_file_location(f"_refjclip.{name}", os.path.join(ref, "jclip", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"_refjclip.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg


def checksum(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.array([float(a.astype(np.float64).sum()), float(np.abs(a).astype(np.float64).sum()), a.size], np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "ref_on_shim.npz"))
    a = ap.parse_args()
    ref = a.reference
    import jclip_b200
    synth = jclip_b200.synth
    torch.manual_seed(0)
    out = {}

    with jt_shim.installed() as jt:
        refpkg = load_ref_jclip(ref)
        nn = sys.modules["jittor.nn"]
        ns = {"jt": jt, "nn": nn, "math": math, "attention": refpkg.mha, "Optional": object, "Tuple": object,
              "softmax": nn.softmax, "dropout": nn.dropout, "np": np}
        lora_names = ["INDEX_POSITIONS_TEXT", "INDEX_POSITIONS_VISION", "LoRALayer", "LinearLoRA", "_canonical_mask",
                      "scaled_dot_product_attention", "_none_or_dtype", "PlainMultiheadAttentionLoRA", "apply_lora"]
        head_names = ["Channel_LP", "logit_normalize", "gaussian_kernel", "cdist", "solve_mta"]
        exec(compile(extract(os.path.join(ref, "test.py"), lora_names + head_names), "reference:test.py", "exec"), ns)
        ood_ns = dict(ns)
        exec(compile(extract(os.path.join(ref, "ood.py"), ["gaussian_kernel", "cdist", "solve_mta"]), "reference:ood.py",
                     "exec"), ood_ns)

        # ---------------- image tower: 2-layer ViT-B/32 geometry (width 768, 12 heads, patch 32, 224 px)
        sd = synth.make_vit_state_dict(seed=21, layers=2)
        imgs = synth.clip_normalize(synth.make_views(22, 1, 3)[0])
        out["tower_images_checksum"] = checksum(imgs)
        out["tower_sd_checksum"] = checksum(np.concatenate([sd[k].ravel() for k in sorted(sd)]))
        model = refpkg.model.build_model({k: jt.array(v) for k, v in sd.items()})
        with torch.no_grad():
            f_zs = model.encode_image(jt.array(imgs))
        out["tower_zero_shot"] = f_zs.numpy()

        # ---------------- LoRA on q,k,v (reference defaults: r=4, alpha=1, dropout 0.25, position all, both towers)
        args = types.SimpleNamespace(encoder="both", position="all", params=["q", "k", "v"], r=4, alpha=1,
                                     dropout_rate=0.25, backbone="ViT-B/32")
        layers = ns["apply_lora"](args, model)
        n_text = len(list(model.transformer.resblocks))
        out["lora_layer_count"] = np.array([len(layers), n_text])
        lora = synth.make_lora(seed=23, layers=2, b_std=0.3)
        for i in range(2):
            layer = layers[n_text + i]                       # text blocks come first (test.py:611-638)
            for name, (A, B) in lora[i].items():
                getattr(layer, name).w_lora_A.data = jt.array(A)
                getattr(layer, name).w_lora_B.data = jt.array(B)
        model.eval()
        with torch.no_grad():
            f_lora = model.encode_image(jt.array(imgs))
        out["tower_lora_qkv"] = f_lora.numpy()
        out["lora_scaling"] = np.array([layers[-1].q_proj.scaling])

        # ---------------- LoRA on q,k,v,o
        model2 = refpkg.model.build_model({k: jt.array(v) for k, v in sd.items()})
        args2 = types.SimpleNamespace(encoder="vision", position="all", params=["q", "k", "v", "o"], r=4, alpha=1,
                                      dropout_rate=0.25, backbone="ViT-B/32")
        layers2 = ns["apply_lora"](args2, model2)
        lora2 = synth.make_lora(seed=24, layers=2, params=("q", "k", "v", "o"), b_std=0.3)
        for i in range(2):
            for name, (A, B) in lora2[i].items():
                getattr(layers2[i], name).w_lora_A.data = jt.array(A)
                getattr(layers2[i], name).w_lora_B.data = jt.array(B)
        model2.eval()
        with torch.no_grad():
            out["tower_lora_qkvo"] = model2.encode_image(jt.array(imgs)).numpy()

        # ---------------- IVLP / VPT image tower of `clip1.load_vlp` (jclip/model1.py): 54 tokens
        sd_vlp = synth.make_vit_state_dict(seed=25, layers=2, vpt_tokens=4)
        out["tower_vlp_sd_checksum"] = checksum(np.concatenate([sd_vlp[k].ravel() for k in sorted(sd_vlp)]))
        design = {"trainer": "IVLP", "vision_depth": 3, "language_depth": 3, "vision_ctx": 4, "language_ctx": 4}
        model_vlp = refpkg.model1.build_model({k: jt.array(v) for k, v in sd_vlp.items()}, design)
        with torch.no_grad():
            out["tower_vlp"] = model_vlp.encode_image(jt.array(imgs)).numpy()

        # ---------------- text tower: CLIP.encode_text (jclip/model.py:202-215), 2 blocks, with and without LoRA
        sd_txt = synth.make_vit_state_dict(seed=27, layers=1, text_layers=2)
        out["text_sd_checksum"] = checksum(np.concatenate([sd_txt[k].ravel() for k in sorted(sd_txt)]))
        tokens = synth.make_tokens(28, 5, vocab=64)
        out["text_tokens"] = tokens
        model_t = refpkg.model.build_model({k: jt.array(v) for k, v in sd_txt.items()})
        with torch.no_grad():
            out["text_zero_shot"] = model_t.encode_text(jt.array(tokens)).numpy()
        args_t = types.SimpleNamespace(encoder="text", position="all", params=["q", "k", "v"], r=4, alpha=1,
                                       dropout_rate=0.25, backbone="ViT-B/32")
        layers_t = ns["apply_lora"](args_t, model_t)
        lora_t = synth.make_lora(seed=29, layers=2, width=512, b_std=0.3)
        for i in range(2):
            for name, (A, B) in lora_t[i].items():
                getattr(layers_t[i], name).w_lora_A.data = jt.array(A)
                getattr(layers_t[i], name).w_lora_B.data = jt.array(B)
        model_t.eval()
        with torch.no_grad():
            out["text_lora_qkv"] = model_t.encode_text(jt.array(tokens)).numpy()

        # ---------------- solve_mta (test.py) and its ood.py twin
        T = synth.make_text_features(seed=31)
        out["mta_text_checksum"] = checksum(T)
        for V in (5, 17, 65):
            X = synth.make_unit_views(40 + V, 2, V)
            out[f"mta_feats_checksum_V{V}"] = checksum(X)
            for img in range(2):
                x, t = jt.array(X[img]), jt.array(T).t()
                # (a) the reference text as is: sqrt of the +-1e-7 diagonal gives NaN for some rows
                jt_shim.CLAMP_SQRT["on"] = False
                with torch.no_grad():
                    m_raw = ns["solve_mta"](x, t)
                # (b) with D^2 clamped at 0 inside jt.sqrt: the oracle's documented definition
                jt_shim.CLAMP_SQRT["on"] = True
                with torch.no_grad():
                    m_clamp = ns["solve_mta"](x, t)
                    lg_clamp = ood_ns["solve_mta"](x, t)
                jt_shim.CLAMP_SQRT["on"] = False
                out[f"mta_mode_raw_V{V}_{img}"] = m_raw.numpy()
                out[f"mta_mode_V{V}_{img}"] = m_clamp.numpy()
                out[f"mta_ood_logits_V{V}_{img}"] = lg_clamp.numpy()

        # ---------------- Channel_LP + logit_normalize (+ the fusion lines of evaluate_base, test.py:1710-1735)
        Tz = synth.make_text_features(seed=33)
        s1, b1, w, b = synth.make_head(34, Tz)
        lp = ns["Channel_LP"]()
        lp.scale1, lp.bias1 = jt.array(s1), jt.array(b1)
        lp.fc.weight, lp.fc.bias = jt.array(w), jt.array(b)
        g = np.random.default_rng(35)
        f = g.standard_normal((4, 512)).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        out["head_feats"] = f
        with torch.no_grad():
            z = lp(jt.array(f))
            out["head_channel_lp"] = z.numpy()
            out["head_logit_normalize_n1"] = ns["logit_normalize"](z[:1]).numpy()
            out["head_logit_normalize_n4"] = ns["logit_normalize"](z).numpy()

    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    np.savez_compressed(a.out, **out)
    print(f"wrote {a.out}: {len(out)} arrays, {os.path.getsize(a.out) / 1024:.1f} KiB")
    for k in ("tower_zero_shot", "tower_lora_qkv", "mta_mode_V65_0"):
        print(k, out[k].shape, float(np.abs(out[k]).mean()))


if __name__ == "__main__":
    main()
