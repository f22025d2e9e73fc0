"""GPU test-time-augmentation view generator: decoded uint8 images -> the [I, 1+N, 3, 224, 224] uint8
view batch `HotPath.evaluate_base` consumes ("next" row f1).

Replaces `JtDataset.__getitem__` (reference test.py:1547-1560), which builds per test image

    transformed_img  = [preprocess(img)]                         # Resize(256, BICUBIC) + CenterCrop(224)   jclip/clip.py:130-135
    transformed_imgs = [transform_s(img) for _ in range(512)]    # RandomResizedCrop(224, scale=(0.2|0.5, 1)) + RandomHorizontalFlip(0.5)
                                                                 # test.py:1898-1903, ood.py:1084-1089

with PIL on the CPU (8 DataLoader workers).  Here the host only draws the crop boxes (a few floats
per view) and uploads the decoded image once; cropping, resampling (Pillow's ImagingResample,
reproduced bit for bit on uint8) and mirroring run in `jcb_tta_views`, so a 65-view batch costs one
image of PCIe traffic instead of 65.  ToTensor's 1/255 and `tfm_clip` stay fused in the tower.

The random stream is numpy's (Jittor's is not reproducible here); boxes are explicit and can be
supplied by the caller for exact replay.
"""
import ctypes
import math

import numpy as np
import torch

from . import _capi
from ._capi import check
from .runtime import get_context, ptr


def centre_view_params(W, H, resize=256, size=224):
    """`Resize(256)` keeps the aspect ratio with the long side truncated by int() (jclip/clip.py:115-124);
    `CenterCrop(224)` rounds its offsets.  Returns (new_w, new_h, left, top)."""
    short, long = (W, H) if W <= H else (H, W)
    if short == resize:
        new_w, new_h = W, H
    else:
        new_short, new_long = resize, int(resize * long / short)
        new_w, new_h = (new_short, new_long) if W <= H else (new_long, new_short)
    return new_w, new_h, int(round((new_w - size) / 2.0)), int(round((new_h - size) / 2.0))


def random_resized_crop_params(rng, W, H, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """`RandomResizedCrop.get_params`: 10 attempts of (area, log-uniform aspect) sampling, then the central-crop
    fallback.  Returns (top, left, h, w)."""
    area = H * W
    lo, hi = math.log(ratio[0]), math.log(ratio[1])
    for _ in range(10):
        target_area = rng.uniform(scale[0], scale[1]) * area
        aspect = math.exp(rng.uniform(lo, hi))
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= W and 0 < h <= H:
            return int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w
    in_ratio = W / H
    if in_ratio < min(ratio):
        w, h = W, int(round(W / min(ratio)))
    elif in_ratio > max(ratio):
        h, w = H, int(round(H * max(ratio)))
    else:
        w, h = W, H
    return (H - h) // 2, (W - w) // 2, h, w


JOB_DTYPE = np.dtype([(n, np.int32) for n in ("image", "top", "left", "crop_h", "crop_w", "out_h", "out_w", "off_y", "off_x",
                                              "filter", "flip", "reserved")])
assert JOB_DTYPE.itemsize == ctypes.sizeof(_capi.ViewJob)


def random_resized_crop_params_batch(rng, W, H, n, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """`n` draws of RandomResizedCrop.get_params for one (W, H), vectorised: the 10 attempts of every draw are
    sampled at once and the first valid one is kept (same distribution as the scalar routine; the order in
    which the stream is consumed differs).  Returns int arrays (top, left, h, w)."""
    area = float(H) * W
    ta = rng.uniform(scale[0], scale[1], (n, 10)) * area
    asp = np.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1]), (n, 10)))
    w = np.rint(np.sqrt(ta * asp)).astype(np.int64)
    h = np.rint(np.sqrt(ta / asp)).astype(np.int64)
    ok = (w > 0) & (w <= W) & (h > 0) & (h <= H)
    first = np.argmax(ok, axis=1)
    any_ok = ok.any(axis=1)
    idx = np.arange(n)
    w, h = w[idx, first], h[idx, first]
    top = np.floor(rng.random(n) * (H - h + 1)).astype(np.int64)
    left = np.floor(rng.random(n) * (W - w + 1)).astype(np.int64)
    if not any_ok.all():       # central-crop fallback
        in_ratio = W / H
        if in_ratio < min(ratio):
            fw, fh = W, int(round(W / min(ratio)))
        elif in_ratio > max(ratio):
            fh, fw = H, int(round(H * max(ratio)))
        else:
            fw, fh = W, H
        bad = ~any_ok
        w[bad], h[bad], top[bad], left[bad] = fw, fh, (H - fh) // 2, (W - fw) // 2
    return top, left, h, w


class TTAViews:
    """images (list of [H, W, 3] uint8 arrays, any sizes) -> torch.uint8 [I, 1 + n_crops, 3, size, size] on the GPU;
    view 0 is the centre view (reference test.py:1700 concatenates it first).

    emit="patches": the same views delivered as the conv1 patch matrix [I, 1 + n_crops, (size/patch)^2, 3 * patch^2] in
    the context's 16-bit operand type (jcb_tta_patches: ToTensor, tfm_clip and the im2col of jclip/model.py:105-108
    fused into the resampler's last pass; bit-identical to the views + the tower's own front end).  `HotPath.evaluate_base`
    takes either form."""

    def __init__(self, n_crops=64, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), size=224, resize=256, flip_p=0.5,
                 seed=0, device=None, emit="views", patch=32, apply_clip_norm=True):
        if emit not in ("views", "patches"):
            raise ValueError(f"emit must be 'views' or 'patches', got {emit!r}")
        self.emit, self.patch, self.apply_clip_norm = emit, patch, apply_clip_norm
        self.n_crops, self.scale, self.ratio = n_crops, scale, ratio
        self.size, self.resize, self.flip_p = size, resize, flip_p
        self.rng = np.random.default_rng(seed)
        self.device = device
        self._pinned = [None, None]     # two grow-only pinned staging buffers for the packed source images ...
        self._copied = [None, None]     # ... and the event that marks the end of the upload out of each
        self._turn = 0

    def draw_jobs(self, shapes):
        """One centre job + n_crops crop jobs per (H, W): a (ViewJob * n) ctypes array."""
        V = 1 + self.n_crops
        jobs = (_capi.ViewJob * (len(shapes) * V))()
        S = self.size
        for i, (H, W) in enumerate(shapes):
            if min(H, W) < 1:
                raise ValueError("empty image")
            new_w, new_h, left, top = centre_view_params(W, H, self.resize, S)
            if new_w < S or new_h < S:
                raise ValueError(f"image {i}: {W}x{H} resizes to {new_w}x{new_h}, smaller than the {S}-pixel centre crop")
            j = jobs[i * V]
            j.image, j.top, j.left, j.crop_h, j.crop_w = i, 0, 0, H, W
            j.out_h, j.out_w, j.off_y, j.off_x = new_h, new_w, top, left
            j.filter, j.flip = _capi.FILTER_BICUBIC, 0
            for c in range(self.n_crops):
                t, l, h, w = random_resized_crop_params(self.rng, W, H, self.scale, self.ratio)
                j = jobs[i * V + 1 + c]
                j.image, j.top, j.left, j.crop_h, j.crop_w = i, t, l, h, w
                j.out_h, j.out_w, j.off_y, j.off_x = S, S, 0, 0
                j.filter, j.flip = _capi.FILTER_BILINEAR, int(self.rng.random() < self.flip_p)
        return jobs

    def draw_jobs_fast(self, shapes):
        """Same jobs as draw_jobs, drawn with numpy vector ops (a structured array of JOB_DTYPE): thousands of
        views per millisecond instead of a Python loop per view."""
        V, S = 1 + self.n_crops, self.size
        jobs = np.zeros(len(shapes) * V, dtype=JOB_DTYPE)
        for i, (H, W) in enumerate(shapes):
            new_w, new_h, left, top = centre_view_params(W, H, self.resize, S)
            if new_w < S or new_h < S:
                raise ValueError(f"image {i}: {W}x{H} resizes to {new_w}x{new_h}, smaller than the {S}-pixel centre crop")
            blk = jobs[i * V:(i + 1) * V]
            blk["image"] = i
            blk[0] = (i, 0, 0, H, W, new_h, new_w, top, left, _capi.FILTER_BICUBIC, 0, 0)
            if self.n_crops:
                t, l, h, w = random_resized_crop_params_batch(self.rng, W, H, self.n_crops, self.scale, self.ratio)
                c = blk[1:]
                c["top"], c["left"], c["crop_h"], c["crop_w"] = t, l, h, w
                c["out_h"] = c["out_w"] = S
                c["filter"] = _capi.FILTER_BILINEAR
                c["flip"] = self.rng.random(self.n_crops) < self.flip_p
        return jobs

    def out_shape_dtype(self, n_images, ctx):
        V, S = 1 + self.n_crops, self.size
        if self.emit == "patches":
            return (n_images, V, (S // self.patch) ** 2, 3 * self.patch * self.patch), ctx.operand_torch_dtype
        return (n_images, V, 3, S, S), torch.uint8

    def __call__(self, images, jobs=None, out=None, stream=None):
        """out: an existing device tensor of out_shape_dtype() to write into (double buffering by the caller);
        stream: a torch.cuda.Stream to upload and generate on instead of the current one (the caller orders it against
        the consumer with events; HotPath.evaluate_image_stream does)."""
        imgs = [np.ascontiguousarray(np.asarray(im), dtype=np.uint8) for im in images]
        for im in imgs:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f"expected [H, W, 3] uint8 RGB images, got {im.shape}")
        shapes = [im.shape[:2] for im in imgs]
        if jobs is None:
            jobs = self.draw_jobs_fast(shapes)
        n_jobs = len(jobs)
        if isinstance(jobs, np.ndarray):
            if jobs.dtype != JOB_DTYPE or not jobs.flags["C_CONTIGUOUS"]:
                raise TypeError("jobs must be a contiguous array of tta.JOB_DTYPE")
            jobs_ptr = jobs.ctypes.data_as(ctypes.POINTER(_capi.ViewJob))
        else:
            jobs_ptr = jobs
        if not torch.cuda.is_available():
            raise RuntimeError("TTAViews runs on a B200 GPU only; there is no CPU fallback")
        ctx = get_context(self.device)
        dev = torch.device("cuda", ctx.device)
        # pack the decoded images into one pinned buffer (16-byte aligned starts) and upload once
        descs = (_capi.SrcImage * len(imgs))()
        off = 0
        for i, im in enumerate(imgs):
            descs[i].offset, descs[i].height, descs[i].width = off, im.shape[0], im.shape[1]
            off += (im.size + 15) // 16 * 16
        # double-buffered: packing batch i+1 on the host overlaps the GPU work of batch i; a buffer is reused only
        # after the upload that last read it has completed (its event), never after a full stream sync
        t = self._turn
        self._turn ^= 1
        if self._pinned[t] is None or self._pinned[t].numel() < off:
            self._pinned[t] = torch.empty(max(off, 16), dtype=torch.uint8, pin_memory=True)
        elif self._copied[t] is not None:
            self._copied[t].synchronize()
        host = self._pinned[t][:max(off, 16)]
        hv = host.numpy()
        for d, im in zip(descs, imgs):
            hv[d.offset:d.offset + im.size] = im.reshape(-1)
        S = self.size
        patches = self.emit == "patches"
        per_view = ((S // self.patch) ** 2, 3 * self.patch * self.patch) if patches else (3, S, S)
        with torch.cuda.device(dev), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
            cur = torch.cuda.current_stream(dev)
            if stream is None:
                ctx.bind_current_stream()
            src = host.to(dev, non_blocking=True)
            self._copied[t] = torch.cuda.Event()
            self._copied[t].record(cur)
            dtype = ctx.operand_torch_dtype if patches else torch.uint8
            want = n_jobs * int(np.prod(per_view))
            if out is None:
                out = torch.empty((n_jobs,) + per_view, dtype=dtype, device=dev)
            elif out.numel() != want or out.dtype != dtype or not out.is_cuda or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous {dtype} device tensor of {want} elements ({n_jobs} views of {per_view})")
            sp = ctypes.c_void_p(cur.cuda_stream)
            if patches:
                check(ctx.lib.jcb_tta_patches(ctx.handle, sp, ptr(src), descs, len(imgs), jobs_ptr, n_jobs, S, self.patch,
                                              int(self.apply_clip_norm),
                                              _capi.OPERAND_F16 if dtype == torch.float16 else _capi.OPERAND_BF16, ptr(out)),
                      ctx.handle)
            else:
                check(ctx.lib.jcb_tta_views(ctx.handle, sp, ptr(src), descs, len(imgs), jobs_ptr, n_jobs, S, ptr(out)), ctx.handle)
        if len(imgs) and n_jobs % len(imgs) == 0:
            return out.view((len(imgs), n_jobs // len(imgs)) + per_view)
        return out.view((n_jobs,) + per_view)


def jobs_to_tuples(jobs):
    """[(image, top, left, crop_h, crop_w, out_h, out_w, off_y, off_x, filter, flip)] for tests / replay."""
    if isinstance(jobs, np.ndarray):
        return [tuple(int(x) for x in j)[:11] for j in jobs.tolist()]
    return [(j.image, j.top, j.left, j.crop_h, j.crop_w, j.out_h, j.out_w, j.off_y, j.off_x, j.filter, j.flip) for j in jobs]
