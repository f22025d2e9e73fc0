"""The thread-block-cluster forms of the tail (ln_post . proj, L2 norm; jclip/model.py:121-124, test.py:1706) and of the
head (test.py:1710-1738), which launch_tail / launch_head pick for calls with few views / images (the reference's own
loop: one image x 65 views per call), must give the SAME BITS as the one-CTA forms used for large batches: a result
must not depend on how a batch is split over calls or ranks (tests/test_gpu_multi.py relies on it).
JCB_TAIL_CLUSTER / JCB_HEAD_CLUSTER = 0 / 1 force one form."""
import os
from ctypes import c_void_p

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = os.environ.get(self.name)
        os.environ[self.name] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop(self.name, None)
        else:
            os.environ[self.name] = self.old


@pytest.mark.parametrize("n_views", [1, 16, 37, 65, 130])
def test_tail_cluster_form_is_bit_identical(jb, cuda_dev, n_views):
    from oracle import vit_encode_image
    sd = jb.synth.make_vit_state_dict(seed=3, layers=1)
    model = jb.jclip.build_model(sd)
    imgs = jb.synth.make_views(9, 1, n_views).reshape(n_views, 3, 224, 224)
    x = torch.from_numpy(imgs).to(cuda_dev)
    outs = {}
    for form in ("0", "1"):
        for normalize in (True, False):
            with _Env("JCB_TAIL_CLUSTER", form):
                outs[form, normalize] = model.visual(x, apply_clip_norm=True, normalize=normalize).cpu()
    for normalize in (True, False):
        assert torch.equal(outs["0", normalize], outs["1", normalize])
    ref = vit_encode_image(sd, imgs, apply_clip_norm=True, normalize=True)
    cos = torch.nn.functional.cosine_similarity(outs["1", True].double(), ref.double(), dim=-1)
    assert cos.min() >= 0.99999
    assert (outs["1", True].norm(dim=-1) - 1).abs().max() < 1e-5
    # the default choice (cluster form for this few views) is one of the two
    assert torch.equal(model.visual(x, apply_clip_norm=True, normalize=True).cpu(), outs["0", True])


def test_tail_cluster_form_text_tower(jb, cuda_dev):
    """Width 512 (k ranges of 64) and the EOT row gather."""
    sd = jb.synth.make_vit_state_dict(seed=6, layers=1, text_layers=2)
    model = jb.jclip.build_model(sd)
    tok = torch.from_numpy(jb.synth.make_tokens(7, 21, vocab=64)).to(cuda_dev)
    with _Env("JCB_TAIL_CLUSTER", "0"):
        a = model.encode_text(tok, normalize=True).cpu()
    with _Env("JCB_TAIL_CLUSTER", "1"):
        b = model.encode_text(tok, normalize=True).cpu()
    assert torch.equal(a, b)


@pytest.mark.parametrize("I", [1, 7, 40])
def test_head_cluster_form_is_bit_identical(jb, cuda_dev, I):
    g = torch.Generator().manual_seed(5)
    unit = lambda t: t / t.norm(dim=-1, keepdim=True)
    modes = [unit(torch.randn(I, 512, generator=g)).to(cuda_dev) for _ in range(3)]
    Ts = [torch.from_numpy(jb.synth.make_text_features(seed=20 + i)).to(cuda_dev) for i in range(3)]
    lp = [torch.from_numpy(a).to(cuda_dev) for a in jb.synth.make_head(2, Ts[2].cpu().numpy())]
    hw = jb._capi.HeadWeights(*[t.data_ptr() for t in lp])
    ctx = jb.get_context(cuda_dev)
    ctx.bind_current_stream()
    res = {}
    for form in ("0", "1"):
        for rank_by in (2, 6):                       # cs1 (what test.py ranks by) and cs5 (the LP++ fusion)
            topk = torch.empty(I, 5, dtype=torch.int32, device=cuda_dev)
            scores = torch.empty(I, 403, device=cuda_dev)
            allsc = torch.empty(I, 7, 403, device=cuda_dev)
            with _Env("JCB_HEAD_CLUSTER", form):
                jb._capi.check(ctx.lib.jcb_head(ctx.handle, *[c_void_p(t.data_ptr()) for t in modes + Ts], jb._capi.byref(hw),
                                                I, 403, 512, rank_by, 5, c_void_p(topk.data_ptr()),
                                                c_void_p(scores.data_ptr()), c_void_p(allsc.data_ptr())), ctx.handle)
                ctx.sync()
            res[form, rank_by] = (topk.cpu(), scores.cpu(), allsc.cpu())
    for rank_by in (2, 6):
        for a, b in zip(res["0", rank_by], res["1", rank_by]):
            assert torch.equal(a, b)
    # sanity against torch: cs1 = 100 m_pt T_pt^T
    ref = 100.0 * modes[0].cpu() @ Ts[0].cpu().T
    assert (res["1", 2][1] - ref).abs().max() <= 1e-3


def test_one_image_pipeline_equals_its_row_of_a_batch(jb, cuda_dev):
    """One image per call (cluster forms) vs the same image inside a batch of 48 (one-CTA forms for the tail)."""
    import types
    sd = jb.synth.make_vit_state_dict(seed=0, layers=2)
    model = jb.jclip.build_model(sd)
    I, V = 48, 17                                      # 816 views: above the tail's cluster threshold
    imgs = torch.from_numpy(jb.synth.make_views(2, I, V)).to(cuda_dev)
    texts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp_np = jb.synth.make_head(2, texts[2].numpy())
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = lp_np
    hp = jb.HotPath(model, jb.TextBank(*texts, cuda_dev), lp, rank_by="cs5")
    top_b, feats_b, sc_b = hp.evaluate_base(imgs, return_feats=True, return_scores=True)
    for i in (0, 17, 47):
        top_1, feats_1, sc_1 = hp.evaluate_base(imgs[i:i + 1], return_feats=True, return_scores=True)
        assert torch.equal(feats_1[0], feats_b[i])
        assert torch.equal(sc_1[0], sc_b[i])
        assert torch.equal(top_1[0], top_b[i])
