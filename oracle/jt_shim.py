"""A torch-backed stand-in for the slice of `jittor` 1.3.8.5 that the reference's hot-path source uses.
TEST INFRASTRUCTURE ONLY -- used by oracle/make_golden.py, in the build container, to execute the
reference's OWN Python source (jclip/model.py, jclip/mha.py and functions cut out of test.py /
ood.py) and record golden vectors.  Never imported by the product, by the GPU tests or by bench.py.

Why: the reference's arithmetic lives in Jittor, which is neither vendored nor installable offline
(SURVEY.md F3/F4).  Running the reference's source on this shim pins what CAN be pinned here --
control flow, operator order, tensor layouts, parameter plumbing (packed-QKV split, LoRA wiring,
early exits of solve_mta) -- against the reference text itself instead of against our reading of it.
What it cannot pin is Jittor's own kernels; each assumed semantic is stated where it is defined
(SURVEY.md Appendix B): LayerNorm eps 1e-5 / biased variance, `jt.argsort` returning (indices,
values), `jt.std` unbiased, `Var.transpose()` with no arguments reversing the dimensions.
"""
import contextlib
import math
import sys
import types

import numpy as np
import torch


class Var(torch.Tensor):
    """torch.Tensor with the handful of Jittor-only call forms the reference uses."""

    def transpose(self, *dims):
        if not dims:                                   # jt.Var.transpose() == reverse all dims
            return self.permute(*reversed(range(self.dim())))
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            return self.permute(*dims[0])
        if len(dims) > 2:
            return self.permute(*dims)
        return super().transpose(*dims)

    def view(self, *shape):                            # jt.Var.view is reshape (no contiguity demand)
        return self.reshape(*shape)

    # Jittor Vars are immutable values: `a += b` rebinds `a` to a new (broadcast) Var
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def argmax(self, dim=None, keepdims=False):        # ASSUMED Jittor semantics: (indices, values)
        t = self.as_subclass(torch.Tensor)
        idx = torch.argmax(t, dim=dim, keepdim=keepdims)
        val = torch.amax(t, dim=dim, keepdim=keepdims)
        return _v(idx), _v(val)

    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))

    def is_training(self):
        return False


def _v(t):
    return t.as_subclass(Var) if isinstance(t, torch.Tensor) else t


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype == torch.float64:
        t = t.float()
    return _v(t)


class Module:
    """jittor.nn.Module: plain-attribute parameters, `execute`, DFS train()/eval() that only flips flags."""

    def __init__(self, *a, **k):
        self._is_train = True

    def __call__(self, *a, **k):
        return self.execute(*a, **k)

    def children(self):
        for k, v in self.__dict__.items():
            if isinstance(v, Module):
                yield k, v

    def modules(self):
        yield self
        for _, c in self.children():
            yield from c.modules()

    def is_training(self):
        return self._is_train

    def train(self, mode=True):
        # Jittor sets `is_train` on every module by DFS and does NOT call the children's train()
        for m in self.modules():
            m._is_train = bool(mode)
        return self

    def eval(self):
        return Module.train(self, False)

    def named_parameters(self, prefix=""):
        out = []
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                out.append((prefix + k, v))
            elif isinstance(v, Module):
                out.extend(v.named_parameters(prefix + k + "."))
        return out

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def state_dict(self):
        return dict(self.named_parameters())

    def load_parameters(self, sd):
        for name, value in sd.items():
            obj = self
            parts = name.split(".")
            ok = True
            for p in parts[:-1]:
                obj = obj[int(p)] if isinstance(obj, Sequential) and p.isdigit() else getattr(obj, p, None)
                if obj is None:
                    ok = False
                    break
            if not ok or not hasattr(obj, parts[-1]):
                continue
            cur = getattr(obj, parts[-1])
            new = _t(value, torch.float32)
            if isinstance(cur, torch.Tensor) and tuple(cur.shape) != tuple(new.shape):
                raise ValueError(f"{name}: shape {tuple(new.shape)} does not match {tuple(cur.shape)}")
            setattr(obj, parts[-1], new)


class Sequential(Module):
    def __init__(self, *mods):
        super().__init__()
        self.layers = list(mods)
        for i, m in enumerate(self.layers):
            self.__dict__[str(i)] = m

    def __iter__(self):
        return iter(self.layers)

    def __len__(self):
        return len(self.layers)

    def __getitem__(self, i):
        return self.layers[i]

    def children(self):
        for i, m in enumerate(self.layers):
            yield str(i), m

    def named_parameters(self, prefix=""):
        out = []
        for i, m in enumerate(self.layers):
            out.extend(m.named_parameters(f"{prefix}{i}."))
        return out

    def execute(self, x):
        for m in self.layers:
            x = m(x)
        return x


class Linear(Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        bound = 1.0 / math.sqrt(in_features)
        self.weight = _v(torch.empty(out_features, in_features).uniform_(-bound, bound))
        self.bias = _v(torch.empty(out_features).uniform_(-bound, bound)) if bias else None

    def execute(self, x):
        return linear(x, self.weight, self.bias)


class LayerNorm(Module):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__()
        n = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.normalized_shape, self.eps = (n,), eps
        self.weight, self.bias = _v(torch.ones(n)), _v(torch.zeros(n))

    def execute(self, x):
        # ASSUMED Jittor semantics: biased variance over the last dim, eps inside the sqrt
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return (x - mean) / torch.sqrt(var + self.eps) * self.weight + self.bias


class Conv2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, **kw):
        super().__init__()
        self.stride, self.padding = stride, padding
        self.weight = _v(torch.randn(out_channels, in_channels, kernel_size, kernel_size) * 0.02)
        self.bias = _v(torch.zeros(out_channels)) if bias else None

    def execute(self, x):
        return _v(torch.nn.functional.conv2d(x, self.weight, self.bias, stride=self.stride, padding=self.padding))


class Embedding(Module):
    def __init__(self, num, dim):
        super().__init__()
        self.weight = _v(torch.randn(num, dim))

    def execute(self, idx):
        return self.weight[idx.long()]


class Dropout(Module):
    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def execute(self, x):
        if self.is_training() and self.p > 0:
            raise RuntimeError("jt_shim: dropout in training mode is outside the inference path")
        return x


def linear(x, weight, bias=None):
    y = x @ weight.t()
    return y + bias if bias is not None else y


def softmax(x, dim=None):
    return _v(torch.softmax(x, dim=dim))


def dropout(x, p=0.5, is_train=False):
    if is_train and p > 0:
        raise RuntimeError("jt_shim: dropout with p > 0 in training mode is outside the inference path")
    return x


def pad(x, p, *a, **k):
    return _v(torch.nn.functional.pad(x, p, *a, **k))


# --- jittor top level ---------------------------------------------------------------------------------
class _Flags:
    use_cuda = 0


CLAMP_SQRT = {"on": False}   # make_golden flips this to evaluate the oracle's D^2 >= 0 definition


def _shape(args):
    if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
        return tuple(args[0])
    return tuple(int(a) for a in args)


def _dt(dtype):
    return torch.float32 if dtype is None else dtype


def zeros(*shape, dtype=None):
    return _v(torch.zeros(_shape(shape), dtype=_dt(dtype)))


def ones(*shape, dtype=None):
    return _v(torch.ones(_shape(shape), dtype=_dt(dtype)))


def empty(*shape, dtype=None):
    return _v(torch.zeros(_shape(shape), dtype=_dt(dtype)))


def randn(*shape, dtype=None):
    return _v(torch.randn(_shape(shape), dtype=_dt(dtype)))


def zeros_like(x, dtype=None):
    return _v(torch.zeros_like(x, dtype=dtype if isinstance(dtype, torch.dtype) else None))


def array(x, dtype=None):
    return _t(x, dtype).clone()


def argsort(x, dim=-1, descending=False):
    # ASSUMED Jittor semantics: returns (indices, sorted values) -- hence `_, sorted_dist = jt.argsort(...)`
    values, idx = torch.sort(x, dim=dim, descending=descending)
    return _v(idx), _v(values)


def jsum(x, dim=None, keepdims=False, keepdim=False):
    kd = keepdims or keepdim
    return _v(torch.sum(x) if dim is None else torch.sum(x, dim=dim, keepdim=kd))


def jmean(x, dim=None, keepdims=False, keepdim=False):
    kd = keepdims or keepdim
    return _v(torch.mean(x) if dim is None else torch.mean(x, dim=dim, keepdim=kd))


def jstd(x):
    # ASSUMED Jittor semantics: global, unbiased (n-1)
    return _v(torch.std(x, unbiased=True))


def jnorm(x, p=2, dim=None, keepdim=False, keepdims=False):
    kd = keepdim or keepdims
    return _v(torch.norm(x, p=p) if dim is None else torch.norm(x, p=p, dim=dim, keepdim=kd))


def jsqrt(x):
    if CLAMP_SQRT["on"]:
        x = torch.clamp(x, min=0.0)
    return _v(torch.sqrt(x))


def concat(xs, dim=0):
    return _v(torch.cat(list(xs), dim=dim))


def _kaiming_uniform_(var, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    t = torch.empty_like(var)
    torch.nn.init.kaiming_uniform_(t, a=a)
    var.data = t
    return var


def _zero_(var):
    var.data = torch.zeros_like(var)
    return var


def _gauss_(var, mean=0.0, std=1.0):
    var.data = torch.randn_like(var) * std + mean
    return var


def _constant_(var, value=0.0):
    var.data = torch.full_like(var, value)
    return var


def _uniformish_(var, *a, **k):
    return var


def build_modules():
    jt = types.ModuleType("jittor")
    nn = types.ModuleType("jittor.nn")
    init = types.ModuleType("jittor.init")
    jt.Var, jt.flags = Var, _Flags()
    jt.float16, jt.float32, jt.float64, jt.float = torch.float16, torch.float32, torch.float64, torch.float32
    jt.bool, jt.int64, jt.int32 = torch.bool, torch.int64, torch.int32
    jt.zeros, jt.ones, jt.empty, jt.randn, jt.zeros_like, jt.array = zeros, ones, empty, randn, zeros_like, array
    jt.argsort, jt.sum, jt.mean, jt.std, jt.norm, jt.sqrt, jt.concat = argsort, jsum, jmean, jstd, jnorm, jsqrt, concat
    jt.exp = lambda x: _v(torch.exp(x))
    jt.sigmoid = lambda x: _v(torch.sigmoid(x))
    jt.matmul = lambda a, b: _v(torch.matmul(a, b))
    jt.bmm = lambda a, b: _v(torch.bmm(a, b))
    jt.logical_not = lambda x: _v(torch.logical_not(x))
    jt.arange = lambda *a, **k: _v(torch.arange(*a, **k))
    jt.triu_ = lambda x, d=0: _v(torch.triu(x, d))
    jt.no_grad = torch.no_grad
    jt.nn, jt.init = nn, init
    nn.Module, nn.Linear, nn.LayerNorm, nn.Conv2d, nn.Sequential, nn.Embedding, nn.Dropout = (
        Module, Linear, LayerNorm, Conv2d, Sequential, Embedding, Dropout)
    nn.softmax, nn.linear, nn.dropout, nn.pad, nn.init = softmax, linear, dropout, pad, init
    init.kaiming_uniform_, init.zero_, init.gauss_, init.constant_ = _kaiming_uniform_, _zero_, _gauss_, _constant_
    init.xavier_uniform_ = init.xavier_gauss_ = _uniformish_
    return {"jittor": jt, "jittor.nn": nn, "jittor.init": init}


@contextlib.contextmanager
def installed():
    """`import jittor` resolves to the shim inside this context (and only there)."""
    mods = build_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield mods["jittor"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
