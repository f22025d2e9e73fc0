"""GPU: the LayerNorm-fold epilogues of the GEMM driven DIRECTLY through the C-ABI (jcb_gemm with the stats / shift /
colsum fields the tower passes between its GEMMs), on residual rows with the statistics of trained CLIP towers that
random-init weights never produce: |row mean| = 20 x the spread (the case where an uncentred 16-bit copy loses 20 x
the precision of bf16(LN(x))) and outlier channels 100 x above the rest.

  producer  EPI_RESID_LNPREP_{SHORT,LONG}:  x += A W^T + b (fp32, in place); copy = round16(x - shift);
            stats[m, n / 256] = (sum, sum of squares) of the centred copy's fp32 values; shift_out = shift
            with shift[m] = shift_in[m] + sum(stats_in[m, :, 0]) / N   (the row mean at the previous LayerNorm point)
  consumer  EPI_LNFOLD_{16,GELU_16}:        out = round16(r (copy Wf^T) - r mu' S + c)  ==  LN(x) W^T + b

Reference: jclip/model.py:17-21 (LayerNorm), :59-62 (x += attn(ln_1 x); x += mlp(ln_2 x)), in fp64 on the CPU.
The bar for the folded path is the stand-alone path built from the same kernels (layernorm kernel -> plain GEMM):
its error may not exceed 1.5 x that one's (+ 1e-4), for bf16 and fp16 operands.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

OPS = [torch.bfloat16, torch.float16]
IDS = ["bf16", "f16"]
ULP = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11}


def _rows(kind, M, W, g):
    """fp32 residual rows [M, W] with unit spread plus the adversarial part."""
    x = torch.randn(M, W, generator=g)
    if kind in ("offset", "both"):
        x = x + 20.0 * (1 + 0.1 * torch.randn(M, 1, generator=g))          # |mean| = 20 std, varying per row
    if kind in ("outliers", "both"):
        x[:, 7] += 100.0
        x[:, 300] -= 100.0
        x[:, 511] += 100.0
    return x


def _stats_of(xc, slots):
    """what the embed kernel leaves for the first block: slot 0 = (sum, sumsq) of the centred row, others 0"""
    st = torch.zeros(xc.shape[0], slots, 2)
    st[:, 0, 0] = xc.sum(-1)
    st[:, 0, 1] = (xc * xc).sum(-1)
    return st


@pytest.mark.parametrize("op", OPS, ids=IDS)
@pytest.mark.parametrize("kind", ["plain", "offset", "outliers", "both"])
@pytest.mark.parametrize("epi_name,K", [("EPI_RESID_LNPREP_SHORT", 768), ("EPI_RESID_LNPREP_LONG", 3072)])
def test_producer_then_consumer(jb, cuda_dev, op, kind, epi_name, K):
    C = jb._capi
    g = torch.Generator().manual_seed({"plain": 1, "offset": 2, "outliers": 3, "both": 4}[kind] * 10 + K % 7)
    M, W, N2 = 1000, 768, 2304                                  # ragged M (not a multiple of 256)
    slots = W // 256
    x0 = _rows(kind, M, W, g)
    # the previous LayerNorm point: shift = a slightly stale mean (as in the tower: the mean one GEMM earlier)
    shift0 = x0.mean(-1) + 0.05 * torch.randn(M, generator=g)
    st0 = _stats_of(x0 - shift0[:, None], slots)
    A = torch.randn(M, K, generator=g).to(op)
    Wp = (torch.randn(W, K, generator=g) * K ** -0.5).to(op)
    bp = 0.1 * torch.randn(W, generator=g)
    d = cuda_dev
    resid = x0.clone().to(d)
    copy = torch.full((M, W), float("nan"), dtype=op, device=d)
    st1 = torch.full((M, slots, 2), float("nan"), device=d)
    shift1 = torch.full((M,), float("nan"), device=d)
    jb.blocks.gemm(A.to(d), Wp.to(d), resid, getattr(C, epi_name), bias=bp.to(d), stats=st1, out2=copy,
                   stats_in=st0.to(d), shift_in=shift0.to(d), shift_out=shift1)
    # ---- producer against fp64
    x1 = x0.double() + A.double() @ Wp.double().t() + bp.double()
    assert (resid.cpu().double() - x1).abs().max() <= 6e-3 * (K / 768) ** 0.5 + 1e-6 * x1.abs().max()
    want_shift = shift0.double() + st0[:, :, 0].double().sum(-1) / W          # == mean(x0) up to rounding
    assert (shift1.cpu().double() - want_shift).abs().max() <= 1e-4
    xc = resid.cpu().double() - shift1.cpu().double()[:, None]                # what the epilogue rounded
    assert ((copy.cpu().double() - xc).abs() <= ULP[op] * xc.abs() + 1e-6).all()
    xc32 = (resid.cpu() - shift1.cpu()[:, None]).double()
    for s in range(slots):
        blk = xc32[:, 256 * s:256 * (s + 1)]
        assert (st1[:, s, 0].cpu().double() - blk.sum(-1)).abs().max() <= 2e-3
        assert ((st1[:, s, 1].cpu().double() - (blk * blk).sum(-1)).abs() <= 1e-5 * (blk * blk).sum(-1) + 1e-3).all()
    # centring works: the copy's magnitude is the row's spread, not its offset
    if kind == "offset":
        assert copy.float().abs().mean() < 2.0

    # ---- consumer: LN(x1) Wc^T + bc through the folded GEMM, against fp64 and against the stand-alone path
    gam = 1 + 0.1 * torch.randn(W, generator=g)
    bet = 0.1 * torch.randn(W, generator=g)
    Wc = torch.randn(N2, W, generator=g) * W ** -0.5
    bc = 0.1 * torch.randn(N2, generator=g)
    Wf, S, c = jb.blocks.fold_ln(Wc.to(d), gam.to(d), bet.to(d), bc.to(d), op)
    out = torch.empty(M, N2, dtype=op, device=d)
    jb.blocks.gemm(copy, Wf, out, C.EPI_LNFOLD_16, bias=c, stats=st1, colsum=S)
    mu = x1.mean(-1, keepdim=True)
    ln = (x1 - mu) / torch.sqrt(((x1 - mu) ** 2).mean(-1, keepdim=True) + 1e-5) * gam.double() + bet.double()
    ref = ln @ Wc.double().t() + bc.double()
    err_fold = (out.cpu().double() - ref).pow(2).mean().sqrt().item()
    # stand-alone: layernorm kernel (fp32 in, 16-bit out) -> plain GEMM with the 16-bit weight
    ln16 = torch.empty(M, W, dtype=op, device=d)
    jb.blocks.layernorm(resid, gam.to(d), bet.to(d), ln16)
    out_sa = torch.empty(M, N2, dtype=op, device=d)
    jb.blocks.gemm(ln16, Wc.to(op).to(d), out_sa, C.EPI_BIAS_16, bias=bc.to(d))
    err_sa = (out_sa.cpu().double() - ref).pow(2).mean().sqrt().item()
    print(f"{kind} {epi_name} {op}: rms error folded {err_fold:.3e}, stand-alone {err_sa:.3e}")
    assert err_fold <= 1.5 * err_sa + 1e-4, (err_fold, err_sa)
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("op", OPS, ids=IDS)
def test_uncentred_copy_is_what_centring_fixes(jb, cuda_dev, op):
    """Documents the failure mode the shift removes: with shift_in = NULL (shift 0, the round-1 behaviour) rows of
    mean 20 x spread lose an order of magnitude of precision through the folded GEMM."""
    C = jb._capi
    g = torch.Generator().manual_seed(3)
    M, W, K, N2 = 512, 768, 768, 768
    slots = W // 256
    d = cuda_dev
    x0 = _rows("offset", M, W, g)
    A = torch.randn(M, K, generator=g).to(op)
    Wp = (torch.randn(W, K, generator=g) * K ** -0.5).to(op)
    gam = 1 + 0.1 * torch.randn(W, generator=g)
    bet = 0.1 * torch.randn(W, generator=g)
    Wc = torch.randn(N2, W, generator=g) * W ** -0.5
    bc = torch.zeros(N2)
    Wf, S, c = jb.blocks.fold_ln(Wc.to(d), gam.to(d), bet.to(d), bc.to(d), op)
    x1 = x0.double() + A.double() @ Wp.double().t()
    mu = x1.mean(-1, keepdim=True)
    ref = ((x1 - mu) / torch.sqrt(((x1 - mu) ** 2).mean(-1, keepdim=True) + 1e-5) * gam.double() + bet.double()) @ Wc.double().t()
    errs = {}
    for centred in (True, False):
        resid = x0.clone().to(d)
        copy = torch.empty(M, W, dtype=op, device=d)
        st1 = torch.empty(M, slots, 2, device=d)
        kw = {}
        if centred:
            shift0 = x0.mean(-1)
            kw = dict(stats_in=_stats_of(x0 - shift0[:, None], slots).to(d), shift_in=shift0.to(d),
                      shift_out=torch.empty(M, device=d))
        jb.blocks.gemm(A.to(d), Wp.to(d), resid, C.EPI_RESID_LNPREP_SHORT, stats=st1, out2=copy, **kw)
        out = torch.empty(M, N2, dtype=op, device=d)
        jb.blocks.gemm(copy, Wf, out, C.EPI_LNFOLD_16, bias=c, stats=st1, colsum=S)
        errs[centred] = (out.cpu().double() - ref).pow(2).mean().sqrt().item()
    print(errs)
    assert errs[False] > 5 * errs[True], errs


def test_lnprep_rejects_aliased_stats(jb, cuda_dev):
    C = jb._capi
    d = cuda_dev
    A = torch.zeros(256, 768, dtype=torch.float16, device=d)
    Wp = torch.zeros(768, 768, dtype=torch.float16, device=d)
    resid = torch.zeros(256, 768, device=d)
    copy = torch.zeros(256, 768, dtype=torch.float16, device=d)
    st = torch.zeros(256, 3, 2, device=d)
    sh = torch.zeros(256, device=d)
    with pytest.raises(jb.JcbError):      # the producer reads stats_in while other tiles write stats: they must differ
        jb.blocks.gemm(A, Wp, resid, C.EPI_RESID_LNPREP_SHORT, stats=st, out2=copy, stats_in=st, shift_in=sh, shift_out=sh)
