"""GPU parity of the building-block kernels, called through the C-ABI (ctypes) exactly as the
product path calls them, for BOTH 16-bit operand types (bf16 and fp16: same kernels, the element type is a
template parameter of the conversions and a field of the tcgen05 instruction descriptor).  References are plain
torch fp32 ops on the SAME 16-bit-rounded operands, so the tolerances below measure accumulation order and output
rounding only.

Tolerances (written here, used below), ulp = 2^-8 relative for bf16 outputs and 2^-11 for fp16 outputs:
  GEMM, fp32 out   :  |d| <= 2e-3 * sqrt(K/768) + 1e-3 * |ref|      (fp32 accumulate, different order)
  GEMM, 16-bit out :  one ulp of the result on top of the above
  LayerNorm        :  one ulp + 1e-3 absolute
  attention        :  P is rounded to 16 bits before P V (as flash-style kernels do): 1e-2 (bf16) / 2e-3 (fp16)
                      absolute + one ulp
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

OPS = [torch.bfloat16, torch.float16]
ULP = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11}
IDS = ["bf16", "f16"]


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


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_f32_bias(jb, cuda_dev, M, N, K, op):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 1.0).to(op).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(op).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    out = torch.full((M, N), float("nan"), device=cuda_dev)
    jb.blocks.gemm(A, B, out, jb._capi.EPI_F32, bias=bias)
    ref = A.float() @ B.float().t() + bias
    tol = 2e-3 * math.sqrt(K / 768) + 1e-3 * ref.abs()
    assert torch.isfinite(out).all()
    assert ((out - ref).abs() <= tol).all(), float((out - ref).abs().max())


def test_gemm_no_bias(jb, cuda_dev):
    g = torch.Generator().manual_seed(5)
    A = torch.randn(300, 768, generator=g).to(torch.bfloat16).to(cuda_dev)
    B = (torch.randn(256, 768, generator=g) * 768 ** -0.5).to(torch.bfloat16).to(cuda_dev)
    out = torch.empty(300, 256, device=cuda_dev)
    jb.blocks.gemm(A, B, out, jb._capi.EPI_F32)
    ref = A.float() @ B.float().t()
    assert (out - ref).abs().max() <= 3e-3


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("epi_name", ["EPI_BIAS_16", "EPI_BIAS_GELU_16"])
def test_gemm_16bit_epilogues(jb, cuda_dev, epi_name, op):
    g = torch.Generator().manual_seed(11)
    M, N, K = 1000, 3072, 768
    A = torch.randn(M, K, generator=g).to(op).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(op).to(cuda_dev)
    bias = (0.1 * torch.randn(N, generator=g)).to(cuda_dev)
    out = torch.empty(M, N, dtype=op, device=cuda_dev)
    jb.blocks.gemm(A, B, out, getattr(jb._capi, epi_name), bias=bias)
    ref = A.float() @ B.float().t() + bias
    if epi_name == "EPI_BIAS_GELU_16":
        ref = ref * torch.sigmoid(1.702 * ref)          # QuickGELU, reference jclip/model.py:27
    # tanh.approx in the GELU epilogue: 2^-11 relative, i.e. one more fp16 ulp
    tol = 3e-3 + 2 * ULP[op] * ref.abs()
    assert ((out.float() - ref).abs() <= tol).all(), float((out.float() - ref).abs().max())


def test_gemm_f16_output_saturates(jb, cuda_dev):
    """fp16 operands: an epilogue result beyond +-65504 is stored as +-65504, never inf."""
    M, N, K = 128, 128, 64
    A = torch.full((M, K), 200.0, dtype=torch.float16, device=cuda_dev)
    B = torch.full((N, K), 100.0, dtype=torch.float16, device=cuda_dev)
    B[1::2] = -100.0
    out = torch.empty(M, N, dtype=torch.float16, device=cuda_dev)
    jb.blocks.gemm(A, B, out, jb._capi.EPI_BIAS_16)
    assert torch.isfinite(out).all()
    assert torch.equal(out[:, 0::2], torch.full((M, N // 2), 65504.0, dtype=torch.float16, device=cuda_dev))
    assert torch.equal(out[:, 1::2], torch.full((M, N // 2), -65504.0, dtype=torch.float16, device=cuda_dev))


@pytest.mark.parametrize("op", OPS, ids=IDS)
def test_gemm_residual_accumulates(jb, cuda_dev, op):
    g = torch.Generator().manual_seed(12)
    M, N, K = 777, 768, 3072
    A = torch.randn(M, K, generator=g).to(op).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * K ** -0.5).to(op).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    resid = torch.randn(M, N, generator=g).to(cuda_dev)
    out = resid.clone()
    jb.blocks.gemm(A, B, out, jb._capi.EPI_BIAS_RESID_F32, bias=bias)
    ref = resid + A.float() @ B.float().t() + bias
    assert (out - ref).abs().max() <= 6e-3


def test_gemm_is_deterministic(jb, cuda_dev):
    g = torch.Generator().manual_seed(13)
    A = torch.randn(5000, 768, generator=g).to(torch.float16).to(cuda_dev)
    B = (torch.randn(2304, 768, generator=g) * 768 ** -0.5).to(torch.float16).to(cuda_dev)
    o1 = torch.empty(5000, 2304, dtype=torch.float16, device=cuda_dev)
    o2 = torch.empty_like(o1)
    jb.blocks.gemm(A, B, o1, jb._capi.EPI_BIAS_16)
    jb.blocks.gemm(A, B, o2, jb._capi.EPI_BIAS_16)
    assert torch.equal(o1, o2)


def test_gemm_rejects_bad_shapes(jb, cuda_dev):
    A = torch.zeros(128, 100, dtype=torch.bfloat16, device=cuda_dev)   # K not a multiple of 64
    B = torch.zeros(128, 100, dtype=torch.bfloat16, device=cuda_dev)
    out = torch.zeros(128, 128, device=cuda_dev)
    with pytest.raises(jb.JcbError):
        jb.blocks.gemm(A, B, out, jb._capi.EPI_F32)


def test_tensor_map_cache_hits(jb, cuda_dev):
    """The second launch on the same (pointer, shape) re-uses the encoded TMA descriptors."""
    import ctypes
    A = torch.zeros(256, 128, dtype=torch.float16, device=cuda_dev)
    B = torch.zeros(128, 128, dtype=torch.float16, device=cuda_dev)
    out = torch.zeros(256, 128, device=cuda_dev)
    lib = jb.load_library()
    h0, m0, h1, m1 = (ctypes.c_uint64() for _ in range(4))
    jb.blocks.gemm(A, B, out, jb._capi.EPI_F32)
    lib.jcb_tensor_map_cache_stats(ctypes.byref(h0), ctypes.byref(m0))
    jb.blocks.gemm(A, B, out, jb._capi.EPI_F32)
    lib.jcb_tensor_map_cache_stats(ctypes.byref(h1), ctypes.byref(m1))
    assert m1.value == m0.value and h1.value >= h0.value + 3


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("rows", [1, 7, 50, 1600, 4001])
def test_layernorm(jb, cuda_dev, rows, op):
    from oracle.vit import layer_norm
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, 768, generator=g) * 2 + 0.3).to(cuda_dev)
    w = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_dev)
    b = (0.1 * torch.randn(768, generator=g)).to(cuda_dev)
    y = torch.empty(rows, 768, dtype=op, device=cuda_dev)
    jb.blocks.layernorm(x, w, b, y)
    ref = layer_norm(x.cpu(), w.cpu(), b.cpu())
    assert ((y.float().cpu() - ref).abs() <= 1e-3 + ULP[op] * ref.abs()).all()


def _attention_ref(qkv, n_views, T, H, causal):
    d = 64
    x = qkv.float().view(n_views, T, 3, H, d)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-2, -1) / math.sqrt(d)                                # jclip/mha.py:55-83
    if causal:                                                                # jclip/model.py:189-193 build_attention_mask
        s = s + torch.triu(torch.full((T, T), float("-inf"), device=s.device), 1)
    return (torch.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).reshape(n_views * T, H * d)


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("n_views,T,H", [(1, 50, 12), (3, 50, 12), (64, 50, 12), (5, 54, 12), (2, 64, 12), (2, 17, 12),
                                         (1, 1, 2), (700, 50, 12), (333, 54, 8), (9, 50, 3)])
def test_attention(jb, cuda_dev, n_views, T, H, op):
    # T <= 64 with an even head count: two heads per 128-row tile (several work items per persistent CTA at
    # 700 x 6 head pairs); H = 3: one head per tile
    g = torch.Generator().manual_seed(n_views * 100 + T)
    W = H * 64
    qkv = (torch.randn(n_views * T, 3 * W, generator=g) * 1.2).to(op).to(cuda_dev)
    out = torch.empty(n_views * T, W, dtype=op, device=cuda_dev)
    jb.blocks.attention(qkv, n_views, T, H, out)
    ref = _attention_ref(qkv, n_views, T, H, False)
    atol = 1e-2 if op == torch.bfloat16 else 2e-3
    assert ((out.float() - ref).abs() <= atol + ULP[op] * ref.abs()).all(), float((out.float() - ref).abs().max())


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("n_seq,T,H,causal", [(1, 77, 8, True), (5, 77, 8, True), (403, 77, 8, True), (3, 65, 8, True),
                                              (2, 128, 4, True), (3, 100, 6, False), (4, 77, 12, False), (2, 33, 3, True)])
def test_attention_one_head_per_tile(jb, cuda_dev, n_seq, T, H, causal, op):
    """Text-tower shape (77 tokens, 8 heads, causal mask) and every other T <= 128 on the tcgen05 kernel's one-head mode."""
    g = torch.Generator().manual_seed(n_seq * 1000 + T)
    W = H * 64
    qkv = (torch.randn(n_seq * T, 3 * W, generator=g) * 1.2).to(op).to(cuda_dev)
    out = torch.full((n_seq * T, W), float("nan"), dtype=op, device=cuda_dev)
    jb.blocks.attention(qkv, n_seq, T, H, out, causal=causal)
    ref = _attention_ref(qkv, n_seq, T, H, causal)
    atol = 1e-2 if op == torch.bfloat16 else 2e-3
    assert torch.isfinite(out.float()).all()
    assert ((out.float() - ref).abs() <= atol + ULP[op] * ref.abs()).all(), float((out.float() - ref).abs().max())


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("dtype", ["u8", "f32", "bf16"])
@pytest.mark.parametrize("norm", [0, 1])
def test_im2col(jb, cuda_dev, dtype, norm, op):
    """tfm_clip + patch extraction (test.py:1301, jclip/model.py:105-108).  uint8 pixels: the 16-bit patches equal the
    reference's fp32 arithmetic ((u / 255) - mean) / std rounded to 16 bits, bit for bit (bf16; fp16 within one ulp:
    the fused multiply-add rounds once where the reference rounds twice); float pixels: one ulp."""
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
    ref = x.view(n, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(n * G * G, 3 * P * P).to(op)
    out = torch.empty(n * G * G, 3 * P * P, dtype=op, device=cuda_dev)
    jb.blocks.im2col(img.to(cuda_dev), R, P, norm, out)
    if dtype == "u8" and op == torch.bfloat16:
        assert torch.equal(out.cpu(), ref)
    else:
        assert ((out.cpu().float() - ref.float()).abs() <= 2 * ULP[op] * ref.float().abs() + 1e-6).all()
