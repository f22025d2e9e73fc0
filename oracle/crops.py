"""Oracle for the TTA view generator ("next" row f1).  TEST INFRASTRUCTURE ONLY.

The reference builds, per test image, 1 centre view + N random crops on the CPU with PIL, inside 8
DataLoader workers (test.py:1547-1560):

  centre view  `preprocess` = jclip/clip.py:130-135  Resize(256, BICUBIC) on the short side, CenterCrop(224), ToTensor
  crops        test.py:1898-1903 / ood.py:1084-1089  T.RandomResizedCrop(224, scale=(0.2|0.5, 1)) [ratio 3/4..4/3,
               BILINEAR], T.RandomHorizontalFlip(0.5), ToTensor

The arithmetic lives in third-party code again: `jittor.transform` (absent) delegates crop / resize to
**Pillow**, which IS installed here (and on the GPU box), so the checker of the GPU kernel is Pillow
itself (`pil_view`), bit for bit on uint8.  `resample_restatement` is a numpy restatement of Pillow's
`ImagingResample` (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, the 8bpc
horizontal / vertical passes with a uint8 intermediate) -- the algorithm the CUDA kernels implement --
and tests/test_crops_cpu.py pins it against Pillow on random boxes.  Pillow version pinned by the
image: 12.2.0.

RandomResizedCrop.get_params follows the published torchvision algorithm that jittor.transform copies
(10 attempts of area / log-ratio sampling, central-crop fallback); the random stream is numpy's, not
Jittor's (not reproducible here), so parity is defined on explicit boxes.
"""
import math

import numpy as np

BILINEAR, BICUBIC = 0, 1
PRECISION_BITS = 32 - 8 - 2


def _filter(kind, x):
    x = np.abs(x)
    if kind == BILINEAR:
        return np.where(x < 1.0, 1.0 - x, 0.0)
    a = -0.5
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def precompute_coeffs(in_size, in0, in1, out_size, kind):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (ksize, xmin[out], n[out], kk[out, ksize] int32)."""
    support0 = 1.0 if kind == BILINEAR else 2.0
    scale = float(np.float32(in1) - np.float32(in0)) / out_size
    filterscale = max(scale, 1.0)
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, np.int32)
    counts = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = _filter(kind, (x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                     # sequential double sum, as in C
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :xmax] = np.trunc(fixed).astype(np.int64).astype(np.int32)
        xmins[xx], counts[xx] = xmin, xmax
    return ksize, xmins, counts, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_restatement(img, out_w, out_h, kind):
    """img [H, W, 3] uint8 -> [out_h, out_w, 3] uint8: Pillow's two-pass 8bpc resample of the whole image."""
    H, W, _ = img.shape
    _, xmin_h, n_h, kk_h = precompute_coeffs(W, 0.0, float(W), out_w, kind)
    _, ymin_v, n_v, kk_v = precompute_coeffs(H, 0.0, float(H), out_h, kind)
    src = img.astype(np.int64)
    tmp = np.empty((H, out_w, 3), np.uint8)
    for xx in range(out_w):
        seg = src[:, xmin_h[xx]:xmin_h[xx] + n_h[xx], :]
        acc = (1 << (PRECISION_BITS - 1)) + np.einsum("hkc,k->hc", seg, kk_h[xx, :n_h[xx]].astype(np.int64))
        tmp[:, xx, :] = _clip8(acc)
    t64 = tmp.astype(np.int64)
    out = np.empty((out_h, out_w, 3), np.uint8)
    for yy in range(out_h):
        seg = t64[ymin_v[yy]:ymin_v[yy] + n_v[yy]]
        acc = (1 << (PRECISION_BITS - 1)) + np.einsum("kwc,k->wc", seg, kk_v[yy, :n_v[yy]].astype(np.int64))
        out[yy] = _clip8(acc)
    return out


# ---- the reference's two transforms, on explicit parameters ----------------------------------------
def centre_view_params(W, H, resize=256, size=224):
    """jclip/clip.py:102-135: Resize(256) keeps the aspect (long side truncated by int()), CenterCrop(224)."""
    short, long = (W, H) if W <= H else (H, W)
    if short == resize:
        new_w, new_h = W, H
    else:
        new_short, new_long = resize, int(resize * long / short)
        new_w, new_h = (new_short, new_long) if W <= H else (new_long, new_short)
    left, top = int(round((new_w - size) / 2.0)), int(round((new_h - size) / 2.0))
    return new_w, new_h, left, top


def random_resized_crop_params(rng, W, H, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """RandomResizedCrop.get_params (torchvision algorithm, copied by jittor.transform) -> (top, left, h, w)."""
    area = H * W
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        target_area = rng.uniform(scale[0], scale[1]) * area
        aspect = math.exp(rng.uniform(log_ratio[0], log_ratio[1]))
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= W and 0 < h <= H:
            top = int(rng.integers(0, H - h + 1))
            left = int(rng.integers(0, W - w + 1))
            return top, left, h, w
    in_ratio = W / H
    if in_ratio < min(ratio):
        w = W
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = H
        w = int(round(h * max(ratio)))
    else:
        w, h = W, H
    return (H - h) // 2, (W - w) // 2, h, w


def pil_centre_view(img, resize=256, size=224):
    """img [H, W, 3] uint8 -> [3, size, size] uint8 through Pillow exactly as `_transform1` does."""
    from PIL import Image
    im = Image.fromarray(img, "RGB")
    new_w, new_h, left, top = centre_view_params(im.size[0], im.size[1], resize, size)
    if (new_w, new_h) != im.size:
        im = im.resize((new_w, new_h), Image.BICUBIC)
    im = im.crop((left, top, left + size, top + size))
    return np.ascontiguousarray(np.asarray(im).transpose(2, 0, 1))


def pil_crop_view(img, top, left, h, w, flip, size=224):
    """RandomResizedCrop's crop_and_resize (BILINEAR) + optional horizontal flip -> [3, size, size] uint8."""
    from PIL import Image
    im = Image.fromarray(img, "RGB").crop((left, top, left + w, top + h)).resize((size, size), Image.BILINEAR)
    if flip:
        im = im.transpose(Image.FLIP_LEFT_RIGHT)
    return np.ascontiguousarray(np.asarray(im).transpose(2, 0, 1))
