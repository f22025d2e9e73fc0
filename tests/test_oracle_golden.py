"""CPU: the oracle against the golden vectors produced by executing the REFERENCE'S OWN SOURCE on the
torch-backed Jittor stand-in (oracle/make_golden.py -> tests/golden/ref_on_shim.npz), plus the
input checksums that guard the seeded generators the vectors were made from."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_on_shim.npz")


def _checksum(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.array([float(a.astype(np.float64).sum()), float(np.abs(a).astype(np.float64).sum()), a.size], np.float64)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def tower_inputs(jb, g):
    sd = jb.synth.make_vit_state_dict(seed=21, layers=2)
    imgs = jb.synth.clip_normalize(jb.synth.make_views(22, 1, 3)[0])
    assert np.allclose(_checksum(imgs), g["tower_images_checksum"], rtol=1e-6)
    assert np.allclose(_checksum(np.concatenate([sd[k].ravel() for k in sorted(sd)])), g["tower_sd_checksum"], rtol=1e-6)
    return sd, imgs


def test_tower_zero_shot(jb, g, tower_inputs):
    from oracle import vit_encode_image
    sd, imgs = tower_inputs
    f = vit_encode_image(sd, imgs).numpy()
    assert f.shape == g["tower_zero_shot"].shape == (3, 512)
    assert np.abs(f - g["tower_zero_shot"]).max() <= 2e-5      # fp32, op order differs (LND vs NLD layout)


def test_tower_lora(jb, g, tower_inputs):
    from oracle import vit_encode_image
    from oracle.vit import lora_scaling
    sd, imgs = tower_inputs
    assert g["lora_scaling"][0] == lora_scaling(4, 1) == 0.5      # alpha / sqrt(r), reference test.py:288-289
    assert g["lora_layer_count"].tolist() == [3, 1]               # 1 text block first, then the 2 vision blocks
    lora = jb.synth.make_lora(seed=23, layers=2, b_std=0.3)
    f = vit_encode_image(sd, imgs, lora=lora, scaling=0.5).numpy()
    assert np.abs(f - g["tower_lora_qkv"]).max() <= 2e-5
    assert np.abs(g["tower_lora_qkv"] - g["tower_zero_shot"]).max() > 0.05   # the adapters matter
    lora2 = jb.synth.make_lora(seed=24, layers=2, params=("q", "k", "v", "o"), b_std=0.3)
    f2 = vit_encode_image(sd, imgs, lora=lora2, scaling=0.5).numpy()
    assert np.abs(f2 - g["tower_lora_qkvo"]).max() <= 2e-5


def test_tower_ivlp_vpt(jb, g, tower_inputs):
    """The 54-token IVLP / VPT tower (reference jclip/model1.py via clip1.load_vlp)."""
    from oracle import vit_encode_image
    _, imgs = tower_inputs
    sd = jb.synth.make_vit_state_dict(seed=25, layers=2, vpt_tokens=4)
    assert np.allclose(_checksum(np.concatenate([sd[k].ravel() for k in sorted(sd)])), g["tower_vlp_sd_checksum"], rtol=1e-6)
    f = vit_encode_image(sd, imgs).numpy()
    assert np.abs(f - g["tower_vlp"]).max() <= 2e-5
    sd0 = {k: v for k, v in sd.items() if k != "visual.VPT"}
    assert np.abs(vit_encode_image(sd0, imgs).numpy() - g["tower_vlp"]).max() > 1e-3     # the prompt tokens matter


def test_merged_equals_applied(jb, tower_inputs):
    """W' = W + s B A (what the device packs) == the reference's un-merged eval math."""
    from oracle import merge_lora_into_state_dict, vit_encode_image
    sd, imgs = tower_inputs
    lora = jb.synth.make_lora(seed=23, layers=2, b_std=0.3)
    a = vit_encode_image(sd, imgs, lora=lora, scaling=0.5)
    b = vit_encode_image(merge_lora_into_state_dict(sd, lora, 0.5), imgs)
    assert (a - b).abs().max() <= 2e-5


@pytest.mark.parametrize("V", [5, 17, 65])
def test_solve_mta(jb, g, V):
    from oracle import solve_mta, solve_mta_logits
    T = torch.from_numpy(jb.synth.make_text_features(seed=31))
    assert np.allclose(_checksum(T.numpy()), g["mta_text_checksum"], rtol=1e-6)
    X = torch.from_numpy(jb.synth.make_unit_views(40 + V, 2, V))
    assert np.allclose(_checksum(X.numpy()), g[f"mta_feats_checksum_V{V}"], rtol=1e-6)
    for i in range(2):
        m = solve_mta(X[i], T.t()).numpy()
        # reference source with D^2 clamped at 0 inside jt.sqrt == the oracle's stated definition
        assert np.abs(m - g[f"mta_mode_V{V}_{i}"]).max() <= 1e-6
        assert np.abs(solve_mta_logits(X[i], T.t()).numpy() - g[f"mta_ood_logits_V{V}_{i}"]).max() <= 1e-4
        # reference source verbatim (NaN self-distances sort last on this backend): bounded deviation
        raw = g[f"mta_mode_raw_V{V}_{i}"]
        assert not np.isnan(raw).any()
        assert np.abs(m - raw).max() <= 1e-3


def test_head(jb, g):
    from oracle import channel_lp, logit_normalize
    Tz = jb.synth.make_text_features(seed=33)
    s1, b1, w, b = (torch.from_numpy(a) for a in jb.synth.make_head(34, Tz))
    f = torch.from_numpy(g["head_feats"])
    z = channel_lp(f, s1, b1, w, b)
    assert np.abs(z.numpy() - g["head_channel_lp"]).max() <= 1e-5
    assert np.abs(logit_normalize(z[:1]).numpy() - g["head_logit_normalize_n1"]).max() <= 1e-5
    assert np.abs(logit_normalize(z).numpy() - g["head_logit_normalize_n4"]).max() <= 1e-5


def test_text_tower(jb, g):
    """CLIP.encode_text of the reference (jclip/model.py:202-215) incl. LoRA on the text blocks."""
    from oracle import text_encode
    sd = jb.synth.make_vit_state_dict(seed=27, layers=1, text_layers=2)
    assert np.allclose(_checksum(np.concatenate([sd[k].ravel() for k in sorted(sd)])), g["text_sd_checksum"], rtol=1e-6)
    tok = jb.synth.make_tokens(28, 5, vocab=64)
    assert np.array_equal(tok, g["text_tokens"])
    assert np.abs(text_encode(sd, tok).numpy() - g["text_zero_shot"]).max() <= 2e-5
    lora = jb.synth.make_lora(seed=29, layers=2, width=512, b_std=0.3)
    assert np.abs(text_encode(sd, tok, lora=lora).numpy() - g["text_lora_qkv"]).max() <= 2e-5
    assert np.abs(g["text_lora_qkv"] - g["text_zero_shot"]).max() > 0.05
