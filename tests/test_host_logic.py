"""CPU: the Python host side that mirrors the reference's API -- build_model shape inference, the LoRA
containers and pickle layout, transforms, tokenizer, result files, sharding arithmetic."""
import gzip
import json
import os
import pickle
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _args(**kw):
    d = dict(encoder="both", position="all", params=["q", "k", "v"], r=4, alpha=1, dropout_rate=0.25, backbone="ViT-B/32")
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_build_model_infers_shapes(jb):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=3, text_layers=2)
    sd["input_resolution"] = np.array(224)          # scalar keys the reference drops (jclip/model.py:280-282)
    m = jb.jclip.build_model(sd)
    assert m.visual.input_resolution == 224 and m.visual.output_dim == 512
    assert len(m.visual.transformer.resblocks) == 3 and len(m.transformer.resblocks) == 2
    assert m.visual.heads == 12 and m.context_length == 77 and m.dtype == torch.float32
    names = dict(m.named_parameters())
    for k in sd:
        if k != "input_resolution":
            assert k in names, k
            assert names[k].shape == sd[k].shape
    assert m.eval() is m and m.cuda() is m and not m.is_train


def test_clip_load_returns_five_tuple(jb, tmp_path):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=1)
    p = tmp_path / "ViT-B-32.pkl"
    with open(p, "wb") as f:
        pickle.dump(sd, f, protocol=4)
    out = jb.clip.load(str(p))
    assert len(out) == 5 and out[0].visual.input_resolution == 224
    assert "ViT-B/32" in jb.clip.available_models()
    with pytest.raises(RuntimeError, match="not found"):
        jb.clip.load("no-such-model")
    from PIL import Image
    img = Image.fromarray((np.random.default_rng(0).random((300, 400, 3)) * 255).astype(np.uint8))
    t1, t2 = out[1](img), out[2](img)
    assert t1.shape == t2.shape == (3, 224, 224) and 0 <= t1.min() and t1.max() <= 1
    assert np.allclose(t2, jb.synth.clip_normalize(t1), atol=1e-6)


def test_apply_lora_layout_and_scaling(jb):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=12, text_layers=12)
    m = jb.jclip.build_model(sd)
    layers = jb.apply_lora(_args(), m)
    assert len(layers) == 24                           # text blocks first, then vision (test.py:611-638)
    assert layers[0].embed_dim == 512 and layers[12].embed_dim == 768
    q = layers[12].q_proj
    assert q.scaling == 0.5 and q.w_lora_A.shape == (4, 768) and q.w_lora_B.shape == (768, 4)
    assert np.all(q.w_lora_B.data == 0) and np.abs(q.w_lora_A.data).max() <= 1 / np.sqrt(768) + 1e-7
    assert not layers[12].proj.lora_enabled            # 'o' not in params
    # packed in_proj rows: q 0:768, k 768:1536, v 1536:2304 (test.py:491-501)
    W = m.visual.transformer.resblocks[0].attn.in_proj_weight.data
    assert np.array_equal(layers[12].k_proj.weight.data, W[768:1536])
    # a second apply_lora finds no plain MultiheadAttention left
    assert jb.apply_lora(_args(), m) == []
    assert jb.apply_lora(_args(encoder="vision", position="up"), jb.jclip.build_model(sd)).__len__() == 4


def test_lora_pickle_schema_matches_shipped_checkpoint(jb, tmp_path):
    """The shipped lora_weights1/lora_weights.pkl (schema recorded in tests/golden/lora_pickle_schema.json):
    a pickle we write has the same nesting, key names, shapes and dtypes, and loads back."""
    schema = json.load(open(os.path.join(GOLD, "lora_pickle_schema.json")))
    sd = jb.synth.make_vit_state_dict(seed=0, layers=12, text_layers=12)
    m = jb.jclip.build_model(sd)
    args = _args(filename="x")
    layers = jb.apply_lora(args, m)
    rng = np.random.default_rng(0)
    for layer in layers:
        for nm in ("q_proj", "k_proj", "v_proj"):
            lin = getattr(layer, nm)
            lin.w_lora_B.data = rng.standard_normal(lin.w_lora_B.shape).astype(np.float32) * 0.006
    path = jb.save_lora(args, 3, layers, save_dir=str(tmp_path))
    assert path.endswith("3_x.pkl")
    d = pickle.load(open(path, "rb"))
    assert d["metadata"] == schema["metadata"]
    assert sorted(d["weights"]) == sorted(schema["layers"])
    for ln, lw in schema["layers"].items():
        assert sorted(d["weights"][ln]) == sorted(lw)
        for pn, pw in lw.items():
            for k, meta in pw.items():
                a = d["weights"][ln][pn][k]
                assert isinstance(a, np.ndarray) and list(a.shape) == meta["shape"] and str(a.dtype) == meta["dtype"]
    m2 = jb.jclip.build_model(sd)
    layers2 = jb.apply_lora(args, m2)
    jb.load_lora(args, layers2, path)
    assert np.array_equal(layers2[17].v_proj.w_lora_B.data, layers[17].v_proj.w_lora_B.data)
    assert m2.visual._dirty


def test_load_lora_errors_match_reference(jb, tmp_path):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=1)
    m = jb.jclip.build_model(sd)
    args = _args(encoder="vision", filename="y")
    layers = jb.apply_lora(args, m)
    with pytest.raises(FileNotFoundError):
        jb.load_lora(args, layers, str(tmp_path / "missing.pkl"))
    path = jb.save_lora(args, 0, layers, save_dir=str(tmp_path))
    for field, bad in (("r", 8), ("alpha", 2), ("encoder", "both"), ("params", ["q", "v"]), ("position", "up")):
        with pytest.raises(ValueError, match="mismatch"):
            jb.load_lora(_args(encoder="vision", **{field: bad}) if field != "encoder" else _args(encoder=bad), layers, path)
    with pytest.raises(ValueError, match="shape mismatch"):
        layers[0].q_proj.w_lora_A.data = np.zeros((8, 768), np.float32)


def test_load_lora_swa_averages(jb, tmp_path):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=1)
    args = _args(encoder="vision", filename="s")
    vals = []
    for e in range(3):
        m = jb.jclip.build_model(sd)
        layers = jb.apply_lora(args, m)
        b = np.full((768, 4), float(e), np.float32)
        layers[0].q_proj.w_lora_B.data = b
        vals.append(layers[0].q_proj.w_lora_A.data.copy())
        jb.save_lora(args, e, layers, save_dir=str(tmp_path))
    m = jb.jclip.build_model(sd)
    layers = jb.apply_lora(args, m)
    jb.load_lora_swa(args, layers, str(tmp_path))
    assert np.allclose(layers[0].q_proj.w_lora_B.data, 1.0)
    assert np.allclose(layers[0].q_proj.w_lora_A.data, np.mean(vals, axis=0), atol=1e-7)


def test_encode_text_runs_with_lora(jb):
    sd = jb.synth.make_vit_state_dict(seed=0, layers=1, text_layers=2)
    m = jb.jclip.build_model(sd)
    tok = torch.randint(1, 60, (3, 77), generator=torch.Generator().manual_seed(0))
    tok[:, 9] = 63
    from oracle import text_encode
    from _torch_text import encode_text_torch
    a = encode_text_torch(m, tok)
    assert (a - text_encode(sd, tok)).abs().max() < 1e-4
    layers = jb.apply_lora(_args(encoder="text"), m)
    assert m._text_dirty
    assert torch.allclose(a, encode_text_torch(m, tok))              # B = 0: adapters are a no-op
    layers[0].q_proj.w_lora_B.data = np.ones((512, 4), np.float32) * 0.05
    assert not torch.allclose(a, encode_text_torch(m, tok), atol=1e-4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.encode_text(tok)


def test_tokenizer_with_a_tiny_vocab(jb, tmp_path, monkeypatch):
    merges = ["#version: 0.2", "h e", "l l", "he ll", "o </w>", "hell o</w>"]
    p = tmp_path / "bpe.txt.gz"
    with gzip.open(p, "wt", encoding="utf-8") as f:
        f.write("\n".join(merges) + "\n")
    from importlib import import_module
    tokmod = import_module("jittor-clip-fewshot_b200.jclip.simple_tokenizer")
    tk = tokmod.SimpleTokenizer(str(p))
    ids = tk.encode("Hello  hello")
    assert len(ids) == 2 and ids[0] == ids[1] == tk.encoder["hello</w>"]
    assert tk.decode(ids).strip() == "hello hello"
    monkeypatch.setenv("JCLIP_BPE_VOCAB", str(p))
    clipmod = import_module("jittor-clip-fewshot_b200.jclip.clip")
    monkeypatch.setattr(clipmod, "_tokenizer", None)
    t = jb.clip.tokenize(["hello", "hello hello"])
    assert t.shape == (2, 77) and t.dtype == torch.int64
    assert t[0, 0] == tk.encoder["<|startoftext|>"] and t[0, 2] == tk.encoder["<|endoftext|>"] and t[0, 3] == 0
    with pytest.raises(RuntimeError, match="too long"):
        jb.clip.tokenize("hello " * 100)
    assert jb.clip.tokenize("hello " * 100, truncate=True)[0, 76] == tk.encoder["<|endoftext|>"]


def test_result_files(jb, tmp_path):
    P = jb.pipeline
    line = P.format_result_line(["TestSetB/img_1.jpg"], [3, 1, 4, 1, 5])
    assert line == "['TestSetB/img_1.jpg'] 3 1 4 1 5"                 # reference test.py:1742
    assert P.process_line(line) == "img_1.jpg 3 1 4 1 5"              # reference test.py:1788-1796
    P.write_results(tmp_path / "r.txt", ["a/b.jpg", "c.jpg"], torch.tensor([[1, 2, 3, 4, 5], [5, 4, 3, 2, 1]]), clean=True)
    assert (tmp_path / "r.txt").read_text() == "b.jpg 1 2 3 4 5\nc.jpg 5 4 3 2 1\n"
    P.write_ood_split(tmp_path / "b.txt", tmp_path / "n.txt", ["x", "y", "z"], torch.tensor([True, False, True]))
    assert (tmp_path / "b.txt").read_text() == "x\nz\n" and (tmp_path / "n.txt").read_text() == "y\n"
    assert P.OOD_BASE_MAX == 372


def test_shard_ranges_cover_exactly(jb):
    for n in (0, 1, 7, 16, 16384, 1000003):
        for world in (1, 2, 3, 8):
            spans = [jb.dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == jb.dist.shard_sizes(n, world)


def test_cls_acc(jb):
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    assert jb.cls_acc(out, torch.tensor([1, 0, 0]), topk=1) == pytest.approx(200 / 3)
    assert jb.cls_acc(out, torch.tensor([1, 0, 1]), topk=2) == pytest.approx(100.0)


def test_u8_normalise_fma_is_exact_in_bf16():
    """im2col fuses ToTensor (u / 255) and tfm_clip ((x - mean) / std, test.py:1301) into one FMA per pixel,
    u * a_c + b_c.  For every one of the 3 x 256 possible inputs the bf16 operand the patch GEMM sees is the one
    the reference's fp32 arithmetic rounds to; likewise u * (1 / 255) without normalisation."""
    mean = np.array([0.48145466, 0.4578275, 0.40821073], np.float32)
    std = np.array([0.26862954, 0.26130258, 0.27577711], np.float32)
    u = np.arange(256, dtype=np.float32)

    def bf16(x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16)

    assert torch.equal(bf16(u * np.float32(1.0 / 255.0)), bf16(u / np.float32(255.0)))
    for c in range(3):
        exact = ((u / np.float32(255.0)) - mean[c]) / std[c]
        a = np.float32(1.0) / (np.float32(255.0) * std[c])
        b = -mean[c] / std[c]
        fma = (u.astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)   # one rounding, as fmaf
        assert torch.equal(bf16(fma), bf16(exact)), c


def test_alias_package_has_single_module_copies():
    """`jclip_b200.x` must be the same module object as the real package's `x` (a second copy would carry its own
    ctypes structure classes and its own context cache)."""
    import importlib
    import jclip_b200
    from jclip_b200.runtime import get_context
    import jclip_b200._capi as capi
    real = importlib.import_module("jittor-clip-fewshot_b200")
    assert jclip_b200 is real
    assert capi is importlib.import_module("jittor-clip-fewshot_b200._capi") and capi is jclip_b200._capi
    assert get_context is real.runtime.get_context
    import jclip_b200.build as b
    assert b is importlib.import_module("jittor-clip-fewshot_b200.build")


def test_result_merger_matches_reference_strings(jb, tmp_path):
    """merge_results / clean_results against what the reference's own update_txt_file + process_line loop produced
    (test.py:1650-1674, :1837-1849; fixtures generated by oracle/make_golden_results.py): string-exact."""
    import json
    import os
    from jclip_b200 import pipeline as P
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "result_files.json")))
    assert len(golden) >= 6
    for name, g in golden.items():
        b, u, r = (tmp_path / f"{name}_{n}.txt" for n in ("base", "update", "result"))
        b.write_text(g["base"])
        u.write_text(g["update"])
        merged = P.merge_results(str(b), str(u))
        assert b.read_text() == g["merged"], name
        assert list(merged) == [ln.split()[0] for ln in g["merged"].splitlines()], name
        P.clean_results(str(b), str(r))
        assert r.read_text() == g["result"], name
    # a blank line is an error in the reference (IndexError on parts[0]); same here
    bad = tmp_path / "bad.txt"
    bad.write_text("a 1 2\n\nb 3 4\n")
    with pytest.raises(IndexError):
        P.read_results(str(bad))


def test_hotpath_default_rank_is_what_test_py_writes(jb):
    """evaluate_base writes topk(cosine_similarity1) (test.py:1738) = "cs1"; unknown scores are refused."""
    import inspect
    from jclip_b200.pipeline import HotPath
    assert inspect.signature(HotPath.__init__).parameters["rank_by"].default == "cs1"
    with pytest.raises(ValueError):
        HotPath(None, None, None, rank_by="cs9")
