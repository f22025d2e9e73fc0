"""GPU parity of the building-block kernels, called through the C-ABI (ctypes) exactly as the
product path calls them.  References are plain torch fp32 ops on the SAME bf16-rounded operands, so
the tolerances below measure accumulation order and output rounding only.

Tolerances (written here, used below):
  GEMM, fp32 out :  |d| <= 2e-3 * sqrt(K/768) + 1e-3 * |ref|      (fp32 accumulate, different order)
  GEMM, bf16 out :  one bf16 ulp of the result (2^-8 relative) on top of the above
  LayerNorm bf16 :  2^-8 relative + 1e-3 absolute
  attention bf16 :  P is rounded to bf16 before P V (as flash-style kernels do): 1e-2 absolute + 2^-8 relative
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctx(jb, dev):
    ctx = jb.get_context(dev)
    ctx.bind_current_stream()
    return ctx


def _gemm(jb, dev, A, B, bias, epi, out, ldo=None):
    from ctypes import c_void_p
    ctx = _ctx(jb, dev)
    M, K = A.shape
    N = B.shape[0]
    jb._capi.check(ctx.lib.jcb_gemm_bf16(ctx.handle, c_void_p(A.data_ptr()), c_void_p(B.data_ptr()), M, N, K,
                                         c_void_p(bias.data_ptr()) if bias is not None else None, epi,
                                         c_void_p(out.data_ptr()), ldo if ldo is not None else N), ctx.handle)
    ctx.sync()


SHAPES = [
    (128, 128, 64),      # one tile, one k-block
    (256, 256, 128),
    (50, 768, 768),      # ragged M < tile
    (4999, 768, 768),    # ragged M, out-proj shape
    (1600, 2304, 768),   # QKV shape (32 views)
    (1600, 3072, 768),   # fc1
    (1600, 768, 3072),   # fc2: 48 k-blocks, many ring wraps
    (1568, 768, 3072),   # patch-embed shape (32 views x 49 patches)
    (6400, 384, 128),    # BN=128 path
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_f32_bias(jb, cuda_dev, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 1.0).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    out = torch.full((M, N), float("nan"), device=cuda_dev)
    _gemm(jb, cuda_dev, A, B, bias, jb._capi.EPI_F32, out)
    ref = A.float() @ B.float().t() + bias
    tol = 2e-3 * math.sqrt(K / 768) + 1e-3 * ref.abs()
    assert torch.isfinite(out).all()
    assert ((out - ref).abs() <= tol).all(), float((out - ref).abs().max())


def test_gemm_no_bias(jb, cuda_dev):
    g = torch.Generator().manual_seed(5)
    A = torch.randn(300, 768, generator=g).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(256, 768, generator=g) * 768 ** -0.5).to(torch.bfloat16).to(cuda_dev)
    out = torch.empty(300, 256, device=cuda_dev)
    _gemm(jb, cuda_dev, A, B, None, jb._capi.EPI_F32, out)
    ref = A.float() @ B.float().t()
    assert (out - ref).abs().max() <= 3e-3


@pytest.mark.parametrize("epi_name", ["EPI_BIAS_BF16", "EPI_BIAS_GELU_BF16"])
def test_gemm_bf16_epilogues(jb, cuda_dev, epi_name):
    g = torch.Generator().manual_seed(11)
    M, N, K = 1000, 3072, 768
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16).to(cuda_dev)
    bias = (0.1 * torch.randn(N, generator=g)).to(cuda_dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=cuda_dev)
    _gemm(jb, cuda_dev, A, B, bias, getattr(jb._capi, epi_name), out)
    ref = A.float() @ B.float().t() + bias
    if epi_name == "EPI_BIAS_GELU_BF16":
        ref = ref * torch.sigmoid(1.702 * ref)          # QuickGELU, reference jclip/model.py:27
    tol = 3e-3 + 2 ** -7 * ref.abs()
    assert ((out.float() - ref).abs() <= tol).all(), float((out.float() - ref).abs().max())


def test_gemm_residual_accumulates(jb, cuda_dev):
    g = torch.Generator().manual_seed(12)
    M, N, K = 777, 768, 3072
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    resid = torch.randn(M, N, generator=g).to(cuda_dev)
    out = resid.clone()
    _gemm(jb, cuda_dev, A, B, bias, jb._capi.EPI_BIAS_RESID_F32, out)
    ref = resid + A.float() @ B.float().t() + bias
    assert (out - ref).abs().max() <= 6e-3


def test_gemm_is_deterministic(jb, cuda_dev):
    g = torch.Generator().manual_seed(13)
    A = torch.randn(5000, 768, generator=g).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(2304, 768, generator=g) * 768 ** -0.5).to(torch.bfloat16).to(cuda_dev)
    o1 = torch.empty(5000, 2304, dtype=torch.bfloat16, device=cuda_dev)
    o2 = torch.empty_like(o1)
    _gemm(jb, cuda_dev, A, B, None, jb._capi.EPI_BIAS_BF16, o1)
    _gemm(jb, cuda_dev, A, B, None, jb._capi.EPI_BIAS_BF16, o2)
    assert torch.equal(o1, o2)


def test_gemm_rejects_bad_shapes(jb, cuda_dev):
    A = torch.zeros(128, 100, dtype=torch.bfloat16, device=cuda_dev)   # K not a multiple of 64
    B = torch.zeros(128, 100, dtype=torch.bfloat16, device=cuda_dev)
    out = torch.zeros(128, 128, device=cuda_dev)
    with pytest.raises(jb.JcbError):
        _gemm(jb, cuda_dev, A, B, None, jb._capi.EPI_F32, out)


@pytest.mark.parametrize("rows", [1, 7, 50, 1600, 4001])
def test_layernorm(jb, cuda_dev, rows):
    from ctypes import c_void_p
    from oracle.vit import layer_norm
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, 768, generator=g) * 2 + 0.3).to(cuda_dev)
    w = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_dev)
    b = (0.1 * torch.randn(768, generator=g)).to(cuda_dev)
    y = torch.empty(rows, 768, dtype=torch.bfloat16, device=cuda_dev)
    ctx = _ctx(jb, cuda_dev)
    jb._capi.check(ctx.lib.jcb_layernorm_bf16(ctx.handle, c_void_p(x.data_ptr()), rows, 768, c_void_p(w.data_ptr()),
                                              c_void_p(b.data_ptr()), c_void_p(y.data_ptr())), ctx.handle)
    ctx.sync()
    ref = layer_norm(x.cpu(), w.cpu(), b.cpu())
    assert ((y.float().cpu() - ref).abs() <= 1e-3 + 2 ** -8 * ref.abs()).all()


@pytest.mark.parametrize("n_views,T,H", [(1, 50, 12), (3, 50, 12), (64, 50, 12), (5, 54, 12), (2, 64, 12), (2, 17, 12),
                                         (1, 1, 2), (700, 50, 12), (333, 54, 8), (9, 50, 3)])
def test_attention(jb, cuda_dev, n_views, T, H):
    # T <= 64 with an even head count runs the tcgen05 / TMEM kernel (several work items per persistent CTA at
    # 700 x 6 head pairs); H = 3 falls back to the mma.sync kernel
    from ctypes import c_void_p
    g = torch.Generator().manual_seed(n_views * 100 + T)
    d = 64
    W = H * d
    qkv = (torch.randn(n_views * T, 3 * W, generator=g) * 1.2).to(torch.bfloat16).to(cuda_dev)
    out = torch.empty(n_views * T, W, dtype=torch.bfloat16, device=cuda_dev)
    ctx = _ctx(jb, cuda_dev)
    jb._capi.check(ctx.lib.jcb_attention_bf16(ctx.handle, c_void_p(qkv.data_ptr()), n_views, T, H,
                                              c_void_p(out.data_ptr())), ctx.handle)
    ctx.sync()
    x = qkv.float().view(n_views, T, 3, H, d)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(d), dim=-1)       # jclip/mha.py:55-83
    ref = (att @ v).permute(0, 2, 1, 3).reshape(n_views * T, W)
    # 1e-2 absolute (P rounded to bf16) + one bf16 ulp of the output itself (2^-8 relative)
    assert ((out.float() - ref).abs() <= 1e-2 + 2.0 ** -8 * ref.abs()).all()


@pytest.mark.parametrize("dtype", ["u8", "f32", "bf16"])
@pytest.mark.parametrize("norm", [0, 1])
def test_im2col(jb, cuda_dev, dtype, norm):
    """tfm_clip + patch extraction (test.py:1301, jclip/model.py:105-108).  uint8 pixels: the bf16 patches equal the
    reference's fp32 arithmetic ((u / 255) - mean) / std rounded to bf16, bit for bit; float pixels: one bf16 ulp."""
    from ctypes import c_void_p
    g = torch.Generator().manual_seed(5)
    n, R, P = 5, 224, 32
    G = R // P
    if dtype == "u8":
        img = torch.randint(0, 256, (n, 3, R, R), generator=g, dtype=torch.uint8)
        x = img.float() / 255.0                                       # T.ToTensor
    else:
        img = torch.rand(n, 3, R, R, generator=g)
        if dtype == "bf16":
            img = img.to(torch.bfloat16)
        x = img.float()
    if norm:
        mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
        std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
        x = (x - mean) / std                                          # T.ImageNormalize
    ref = x.view(n, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(n * G * G, 3 * P * P).to(torch.bfloat16)
    out = torch.empty(n * G * G, 3 * P * P, dtype=torch.bfloat16, device=cuda_dev)
    ctx = _ctx(jb, cuda_dev)
    d = img.to(cuda_dev)
    jb._capi.check(ctx.lib.jcb_im2col_bf16(ctx.handle, c_void_p(d.data_ptr()), {"f32": 0, "bf16": 1, "u8": 2}[dtype], n, R, P,
                                           norm, c_void_p(out.data_ptr())), ctx.handle)
    ctx.sync()
    if dtype == "u8":
        assert torch.equal(out.cpu(), ref)
    else:
        assert ((out.cpu().float() - ref.float()).abs() <= 2.0 ** -7 * ref.float().abs() + 1e-6).all()   # one bf16 ulp
