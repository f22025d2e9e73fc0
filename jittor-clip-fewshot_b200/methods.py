"""Drop-in replacements for the free functions of the hot path in the reference scripts.

  solve_mta(image_features, text_features)      test.py:1391-1461   -> mode [1, 512]
  solve_mta_logits(image_features, text)        ood.py:751-820      -> 100 * mode @ text  [1, C]
  Channel_LP                                    test.py:1223-1234
  logit_normalize(logit)                        test.py:1304-1308
  cls_acc(output, target, topk)                 test.py:821-826

Same names, argument order and shapes as the reference; each is one call into libjclip_b200.so.
Inputs may be torch CUDA tensors or anything exporting DLPack (jittor Vars); outputs are torch CUDA
tensors on the input's device.  Batched variants (`*_batched`) take a leading image dimension, which
is how the library is meant to be driven: one launch for a whole shard of images.
"""
import numpy as np
import torch

from . import _capi
from ._capi import check
from .jclip.model import Module, Param
from .runtime import as_torch, dev_f32, get_context, ptr


def _cuda(x):
    t = as_torch(x)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("jclip_b200 methods run on a B200 GPU only; there is no CPU fallback")
        t = t.cuda()
    return t.to(torch.float32).contiguous()


def solve_mta_batched(image_features, text_features, return_logits=False, params=None):
    """image_features [I, V, D] unit rows (view 0 = un-augmented), text_features [D, C] (i.e. `text.t()`,
    the orientation the reference passes).  Returns modes [I, D] (and 100 * mode @ text [I, C])."""
    x = _cuda(image_features)
    t = dev_f32(text_features, x.device)
    if x.dim() != 3 or t.dim() != 2 or t.shape[0] != x.shape[2]:
        raise ValueError(f"expected feats [I,V,D] and text [D,C]; got {tuple(x.shape)} and {tuple(t.shape)}")
    I, V, D = x.shape
    C = t.shape[1]
    with torch.cuda.device(x.device):
        ctx = get_context(x.device)
        ctx.bind_current_stream()
        mode = torch.empty((I, D), dtype=torch.float32, device=x.device)
        logits = torch.empty((I, C), dtype=torch.float32, device=x.device) if return_logits else None
        p = None
        if params:
            p = _capi.MtaParams()
            ctx.lib.jcb_mta_default_params(_capi.byref(p))
            for k, v in params.items():
                setattr(p, k, v)
        check(ctx.lib.jcb_mta(ctx.handle, ptr(x), ptr(t), I, V, C, D, _capi.byref(p) if p is not None else None,
                              ptr(mode), ptr(logits)), ctx.handle)
    return (mode, logits) if return_logits else mode


def solve_mta(image_features, text_features):
    """reference test.py:1391-1461: [V,512] x [512,C] -> mode [1,512]."""
    x = _cuda(image_features)
    return solve_mta_batched(x.unsqueeze(0), text_features)


def solve_mta_logits(image_features, text_features):
    """reference ood.py:751-820 (`solve_mta` there returns the logits of the mode): -> [1, C]."""
    x = _cuda(image_features)
    return solve_mta_batched(x.unsqueeze(0), text_features, return_logits=True)[1]


class _Fc(Module):
    def __init__(self, out_features, in_features, rng):
        super().__init__()
        bound = 1.0 / np.sqrt(in_features)
        self.weight = Param(rng.uniform(-bound, bound, (out_features, in_features)).astype(np.float32))
        self.bias = Param(rng.uniform(-bound, bound, (out_features,)).astype(np.float32))


class Channel_LP(Module):
    """reference test.py:1223-1234: fc(scale1 * features + bias1), fc = Linear(512, 403)."""

    def __init__(self, dim=512, num_classes=403, seed=0):
        super().__init__()
        self.scale1 = Param(np.ones(dim, np.float32))
        self.bias1 = Param(np.zeros(dim, np.float32))
        self.fc = _Fc(num_classes, dim, np.random.default_rng(seed))
        self._dev = {}

    def device_weights(self, device):
        """(scale1, bias1, fc.weight, fc.bias) as device tensors, re-uploaded when a Param changed."""
        key = (str(device), self.scale1._version, self.bias1._version, self.fc.weight._version, self.fc.bias._version)
        if self._dev.get("key") != key:
            self._dev = {"key": key, "t": tuple(dev_f32(p.data, device) for p in
                                                (self.scale1, self.bias1, self.fc.weight, self.fc.bias))}
        return self._dev["t"]

    def head_struct(self, device):
        s, b, w, fb = self.device_weights(device)
        return _capi.HeadWeights(s.data_ptr(), b.data_ptr(), w.data_ptr(), fb.data_ptr())

    def execute(self, features):
        x = _cuda(features)
        flat = x.reshape(-1, x.shape[-1])
        w = self.device_weights(x.device)
        C, D = w[2].shape
        with torch.cuda.device(x.device):
            ctx = get_context(x.device)
            ctx.bind_current_stream()
            out = torch.empty((flat.shape[0], C), dtype=torch.float32, device=x.device)
            hs = self.head_struct(x.device)
            check(ctx.lib.jcb_channel_lp(ctx.handle, ptr(flat), flat.shape[0], C, D, _capi.byref(hs), ptr(out)), ctx.handle)
        return out.reshape(*x.shape[:-1], C)

    def load(self, path):
        import pickle
        with open(path, "rb") as f:
            sd = pickle.load(f)
        for k, p in self.named_parameters():
            if k in sd:
                p.data = sd[k]


def logit_normalize(logit):
    """reference test.py:1304-1308: (logit - rowmean) / std(all entries, unbiased)."""
    x = _cuda(logit)
    if x.dim() != 2:
        raise ValueError("logit_normalize expects [n, C]")
    with torch.cuda.device(x.device):
        ctx = get_context(x.device)
        ctx.bind_current_stream()
        out = torch.empty_like(x)
        check(ctx.lib.jcb_logit_normalize(ctx.handle, ptr(x), x.shape[0], x.shape[1], ptr(out)), ctx.handle)
    return out


def cosine_topk(features, text_features, k=5, scale=100.0, return_scores=False):
    """`(scale * f @ T.t()).topk(k)` (evaluate_new test.py:1770-1774; k=1: OOD argmax ood.py:875-877).
    features [n, D], text_features [C, D] -> int32 [n, k] (ties: lowest index first)."""
    x = _cuda(features)
    t = dev_f32(text_features, x.device)
    n, D = x.shape
    C = t.shape[0]
    with torch.cuda.device(x.device):
        ctx = get_context(x.device)
        ctx.bind_current_stream()
        idx = torch.empty((n, k), dtype=torch.int32, device=x.device)
        sc = torch.empty((n, C), dtype=torch.float32, device=x.device) if return_scores else None
        check(ctx.lib.jcb_cosine_topk(ctx.handle, ptr(x), ptr(t), n, C, D, float(scale), k, ptr(idx), ptr(sc)), ctx.handle)
    return (idx, sc) if return_scores else idx


def clip_classifier(templates_dict, clip_model, tokenize=None):
    """reference test.py:920-940: per class, tokenize every template, encode_text, L2-normalise, average,
    re-normalise; returns [1, C, D] (callers `.squeeze(0)`).  One batched `encode_text` call for all templates
    instead of one call per template; `tokenize` defaults to `jclip.clip.tokenize`; `templates_dict` values may
    also be pre-tokenised int64 arrays [T_c, context]."""
    if tokenize is None:
        from .jclip.clip import tokenize
    toks, offsets = [], [0]
    for _, templates in templates_dict.items():
        t = templates if isinstance(templates, (np.ndarray, torch.Tensor)) else tokenize(list(templates))
        t = as_torch(t).to(torch.int64).reshape(-1, as_torch(t).shape[-1])
        toks.append(t.cpu())
        offsets.append(offsets[-1] + t.shape[0])
    emb = clip_model.encode_text(torch.cat(toks, dim=0), normalize=True)      # test.py:927-929
    C, D = len(offsets) - 1, emb.shape[1]
    with torch.cuda.device(emb.device):
        ctx = get_context(emb.device)
        ctx.bind_current_stream()
        off = torch.tensor(offsets, dtype=torch.int32, device=emb.device)
        out = torch.empty((C, D), dtype=torch.float32, device=emb.device)
        check(ctx.lib.jcb_class_mean(ctx.handle, ptr(emb), ptr(off), C, D, ptr(out)), ctx.handle)
    return out.unsqueeze(0)


def cls_acc(output, target, topk=1):
    """reference test.py:821-826 (bookkeeping on the host; not a kernel)."""
    out = as_torch(output).detach().float().cpu()
    tgt = as_torch(target).detach().cpu().view(1, -1)
    pred = out.topk(topk, 1, True, True)[1].t()
    correct = pred.eq(tgt.expand_as(pred))
    return 100.0 * float(correct[:topk].reshape(-1).float().sum()) / tgt.shape[1]
