"""`jclip.clip` API surface (reference jclip/clip.py): `load`, `tokenize`, `available_models`.

`load(name)` returns the same 5-tuple as the reference (jclip/clip.py:170-187):
    (model, _transform1, _transform2, tfm_train_base, tfm_train_base1)
where `model.encode_image` runs on the sm_100a library.  Checkpoints are pickles of
{state-dict key: array} (what `jt.load` reads; plain `pickle` suffices); there is no
network here, so the URL table is kept for `available_models()` only and `load` of a model *name*
raises unless the file is already in `download_root`.
"""
import os
import pickle
from typing import List, Union

import numpy as np
import torch

from .model import build_model

__all__ = ["available_models", "load", "load_vlp", "tokenize"]

_MODELS = {   # names of reference jclip/clip.py:19-38 -> checkpoint file name
    "RN50": "RN50.pt", "RN101": "RN101.pt", "RN50x4": "RN50x4.pt", "RN50x16": "RN50x16.pt", "RN50x64": "RN50x64.pt",
    "ViT-B/32": "ViT-B-32.pt", "ViT-B/16": "ViT-B-16.pt", "ViT-L/14": "ViT-L-14.pt",
    "ViT-L/14@336px": "ViT-L-14-336px.pt",
}
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def available_models() -> List[str]:
    return list(_MODELS.keys())


# ---- PIL transforms (host side; the GPU crop generator is the f1 "next" row) -------------------------
class Compose:
    def __init__(self, ts):
        self.transforms = list(ts)

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


class Resize:
    """Short side to `size`, aspect kept (reference jclip/clip.py:102-127)."""

    def __init__(self, size, mode=None):
        from PIL import Image
        self.size, self.mode = size, Image.BICUBIC if mode is None else mode

    def __call__(self, img):
        w, h = img.size
        short, long = (w, h) if w <= h else (h, w)
        if short == self.size:
            return img
        new_short, new_long = self.size, int(self.size * long / short)
        new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
        return img.resize((new_w, new_h), self.mode)


class CenterCrop:
    def __init__(self, size):
        self.size = size

    def __call__(self, img):
        w, h = img.size
        left, top = int(round((w - self.size) / 2.0)), int(round((h - self.size) / 2.0))
        return img.crop((left, top, left + self.size, top + self.size))


class RandomHorizontalFlip:
    def __init__(self, p=0.5, rng=None):
        self.p, self.rng = p, rng or np.random.default_rng()

    def __call__(self, img):
        from PIL import Image
        return img.transpose(Image.FLIP_LEFT_RIGHT) if self.rng.random() < self.p else img


class ToTensor:
    """PIL / HWC uint8 -> CHW float32 in [0,1] (numpy), like jittor.transform.ToTensor."""

    def __call__(self, img):
        a = np.asarray(img)
        if a.ndim == 2:
            a = a[:, :, None]
        if a.dtype == np.uint8:
            a = a.astype(np.float32) / 255.0
        return np.ascontiguousarray(a.transpose(2, 0, 1), dtype=np.float32) if a.shape[-1] in (1, 3, 4) else a


class ImageNormalize:
    """Works on PIL images (-> CHW float32) and on CHW / NCHW arrays or tensors (test.py:1301 applies it to a batch)."""

    def __init__(self, mean, std):
        self.mean = np.asarray(mean, np.float32).reshape(-1, 1, 1)
        self.std = np.asarray(std, np.float32).reshape(-1, 1, 1)

    def __call__(self, img):
        if isinstance(img, torch.Tensor):
            m = torch.from_numpy(self.mean).to(img.device)
            s = torch.from_numpy(self.std).to(img.device)
            return (img - m) / s
        if not isinstance(img, np.ndarray):
            img = ToTensor()(img)
        return (img - self.mean) / self.std


def _transform1(n_px):
    return Compose([Resize(256), CenterCrop(224), ToTensor()])


def _transform2(n_px):
    return Compose([Resize(256), CenterCrop(224), ImageNormalize(CLIP_MEAN, CLIP_STD), ToTensor()])


def tfm_train_base(n_px):
    return Compose([RandomHorizontalFlip(p=0.5), Resize(256), CenterCrop(224), ToTensor()])


def tfm_train_base1(n_px):
    return Compose([RandomHorizontalFlip(p=0.5), Resize(256), CenterCrop(224), ImageNormalize(CLIP_MEAN, CLIP_STD),
                    ToTensor()])


def load_state_dict(path):
    """Read a checkpoint written by `jt.save` / `pickle.dump`: {key: array-like}."""
    with open(path, "rb") as f:
        sd = pickle.load(f)
    if not isinstance(sd, dict):
        raise RuntimeError(f"{path}: expected a pickled state dict, got {type(sd).__name__}")
    return sd


IVLP_DESIGN = {"trainer": "IVLP", "vision_depth": 3, "language_depth": 3, "vision_ctx": 4, "language_ctx": 4}


def load_vlp(name, download_root=None, mode='vit'):
    """reference jclip/clip1.py:189-213: the IVLP / VPT model (54-token image tower)."""
    return load(name, download_root, mode, design_details=dict(IVLP_DESIGN))


def load(name, download_root=None, mode='vit', design_details=None):
    """reference jclip/clip.py:170-187."""
    if name in _MODELS:
        root = download_root or os.path.expanduser("~/.cache/clip")
        model_path = os.path.join(root, _MODELS[name])
        if not os.path.isfile(model_path):
            raise RuntimeError(f"Model {name}: {model_path} is not present and this environment has no network; "
                               f"pass a local checkpoint path instead")
    elif os.path.isfile(name):
        model_path = name
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if mode != 'vit':
        raise NotImplementedError("mode != 'vit' (ModifiedResNet, jclip/model_res.py) is outside the hot path")
    model = build_model(load_state_dict(model_path), design_details)
    n_px = model.visual.input_resolution
    return model, _transform1(n_px), _transform2(n_px), tfm_train_base(n_px), tfm_train_base1(n_px)


_tokenizer = None


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False):
    """reference jclip/clip.py:190-214 -> int64 [n, context_length] (torch, host)."""
    global _tokenizer
    if _tokenizer is None:
        from .simple_tokenizer import SimpleTokenizer
        _tokenizer = SimpleTokenizer()
    if isinstance(texts, str):
        texts = [texts]
    sot, eot = _tokenizer.encoder["<|startoftext|>"], _tokenizer.encoder["<|endoftext|>"]
    result = torch.zeros((len(texts), context_length), dtype=torch.int64)
    for i, text in enumerate(texts):
        tokens = [sot] + _tokenizer.encode(text) + [eot]
        if len(tokens) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
            tokens = tokens[:context_length]
            tokens[-1] = eot
        result[i, :len(tokens)] = torch.tensor(tokens, dtype=torch.int64)
    return result
