"""GPU: the tcgen05/TMEM/TMA bf16 GEMM (sc_linear_bf16) against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands.  Tolerance: fp32 accumulation order only (2e-3 relative to
the row scale), bf16 output additionally rounds to 8 bits of mantissa."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # (M, N, K, relu, residual, bf16_out)
    (128, 128, 64, 0, False, False),
    (128, 128, 256, 0, False, False),
    (256, 256, 256, 0, False, False),
    (300, 768, 256, 0, False, False),       # ragged M, QKV shape
    (2560, 256, 256, 0, True, False),       # decoder projections with residual
    (2560, 2048, 256, 1, False, True),      # FFN1: ReLU + bf16 output
    (2560, 256, 2048, 0, True, False),      # FFN2
    (10752, 2048, 256, 1, False, True),     # encoder FFN1 at 256 streams x 1 block
    (10752, 256, 2048, 0, True, False),
    (77, 1024, 256, 0, False, False),       # output layer, small M
    (640, 512, 256, 0, False, False),       # cross K|V
    (3000, 256, 4864, 0, False, False),     # embed.out
]


@pytest.mark.parametrize("M,N,K,relu,res,bf16_out", SHAPES)
def test_linear_bf16_matches_fp32_reference(M, N, K, relu, res, bf16_out):
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g, device="cuda")).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    r = torch.randn(M, N, generator=g, device="cuda")
    y = r.clone() if res else torch.full((M, N), float("nan"), device="cuda")
    y16 = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") if bf16_out else None
    _lib.check(lib.sc_linear_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr() if res else None,
                                  y.data_ptr(), y16.data_ptr() if bf16_out else None, M, N, K, relu, None), "linear_bf16")
    torch.cuda.synchronize()
    want = a.float() @ w.float().t() + bias
    if relu:
        want = want.relu()
    if res:
        want = want + r
    err = (y - want).abs().max().item()
    scale = want.abs().max().item()
    assert np.isfinite(err) and err <= 2e-3 * max(1.0, scale), f"max err {err} (scale {scale})"
    if bf16_out:
        err16 = (y16.float() - want).abs().max().item()
        assert err16 <= 1e-2 * max(1.0, scale)


@pytest.mark.parametrize("M,K", [(77, 256), (2560, 256), (1000, 2048), (10752, 256)])
def test_linear_bf16_fused_layernorm(M, K):
    """GEMM + bias + in-place residual with the next LayerNorm fused into the epilogue (BN = N = 256)."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(256, K, generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(256, generator=g, device="cuda")
    r = torch.randn(M, 256, generator=g, device="cuda")
    lw = 1.0 + 0.1 * torch.randn(256, generator=g, device="cuda")
    lb = 0.1 * torch.randn(256, generator=g, device="cuda")
    y = r.clone()
    ln = torch.zeros(M, 256, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.sc_linear_bf16_ln(a.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), y.data_ptr(), lw.data_ptr(),
                                     lb.data_ptr(), ln.data_ptr(), M, K, None), "linear_bf16_ln")
    torch.cuda.synchronize()
    want = a.float() @ w.float().t() + bias + r
    assert (y - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    want_ln = torch.nn.functional.layer_norm(want, (256,), lw, lb, 1e-12)
    assert (ln.float() - want_ln).abs().max().item() <= 3e-2


@pytest.mark.parametrize("M,N,relu,bf16_out", [(77, 768, 0, False), (2560, 256, 0, False), (2340, 2048, 1, True),
                                                (10752, 768, 0, True), (300, 1024, 0, False)])
def test_linear_bf16_layernorm_prologue(M, N, relu, bf16_out):
    """LayerNorm computed inside the GEMM (A tile written in the swizzled operand layout by the epilogue warps):
    must equal the unfused LayerNorm(bf16) -> tensor-core GEMM pair."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = 2.0 * torch.randn(M, 256, generator=g, device="cuda") + 0.3
    lw = 1.0 + 0.1 * torch.randn(256, generator=g, device="cuda")
    lb = 0.1 * torch.randn(256, generator=g, device="cuda")
    w = (torch.randn(N, 256, generator=g, device="cuda") / 16).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    y = torch.full((M, N), float("nan"), device="cuda")
    y16 = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") if bf16_out else None
    _lib.check(lib.sc_linear_bf16_lnA(x.data_ptr(), lw.data_ptr(), lb.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                      y.data_ptr(), y16.data_ptr() if bf16_out else None, M, N, relu, None), "lnA")
    torch.cuda.synchronize()
    a = torch.nn.functional.layer_norm(x, (256,), lw, lb, 1e-12).to(torch.bfloat16)
    want = a.float() @ w.float().t() + bias
    if relu:
        want = want.relu()
    # a one-ulp difference in a bf16 A element (LayerNorm rounding) moves an output by at most ~|a| * |w| * 2^-8
    assert (y - want).abs().max().item() <= 2e-2
    assert (y - want).abs().mean().item() <= 1e-3
    if bf16_out:
        assert (y16.float() - want).abs().max().item() <= 5e-2


@pytest.mark.parametrize("M,F", [(128, 128), (77, 256), (300, 2048), (2688, 2048), (10752, 2048)])
def test_ffn_fused_matches_two_gemm_path(M, F):
    """Fused FFN1 -> ReLU -> FFN2 kernel (hidden kept on the SM) against (a) the two tcgen05 GEMMs it replaces, with
    the same bf16 rounding of the hidden activation: same accumulation order, so equal to fp32 rounding noise; and
    (b) a plain PyTorch fp32 reference of the op (tolerance: bf16 rounding of the hidden activation)."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    D = 256
    g = torch.Generator(device="cuda").manual_seed(M * 5 + F)
    a = torch.randn(M, D, generator=g, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(F, D, generator=g, device="cuda") / D ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(F, generator=g, device="cuda") * 0.1
    w2 = (torch.randn(D, F, generator=g, device="cuda") / F ** 0.5).to(torch.bfloat16)
    b2 = torch.randn(D, generator=g, device="cuda") * 0.1
    r = torch.randn(M, D, generator=g, device="cuda")
    # two-GEMM path
    h16 = torch.zeros(M, F, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.sc_linear_bf16(a.data_ptr(), w1.data_ptr(), b1.data_ptr(), None, None, h16.data_ptr(), M, F, D, 1, None), "ffn1")
    y2 = r.clone()
    _lib.check(lib.sc_linear_bf16(h16.data_ptr(), w2.data_ptr(), b2.data_ptr(), y2.data_ptr(), y2.data_ptr(), None, M, D, F, 0, None), "ffn2")
    # fused, in place on the residual like the engine (accumulate), and as a plain store
    y = r.clone()
    _lib.check(lib.sc_ffn_bf16(a.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), y.data_ptr(),
                               1, M, F, 1, None), "ffn_fused")
    y0 = torch.full((M, D), float("nan"), device="cuda")
    _lib.check(lib.sc_ffn_bf16(a.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), y0.data_ptr(),
                               0, M, F, 1, None), "ffn_fused")
    # hidden dimension split over several CTAs per tile, partial tiles added at the L2 (order-dependent rounding)
    ys = None
    if F >= 512:
        ys = r.clone()
        _lib.check(lib.sc_ffn_bf16(a.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), ys.data_ptr(),
                                   1, M, F, 4, None), "ffn_fused_split")
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(y0).all()
    d = (y - y2).abs().max().item()
    assert d <= 1e-5 * max(1.0, y2.abs().max().item()), f"fused vs two-GEMM: {d}"
    h = (a.float() @ w1.float().t() + b1).relu()
    want = h.to(torch.bfloat16).float() @ w2.float().t() + b2 + r
    err = (y - want).abs().max().item()
    assert err <= 5e-3 * max(1.0, want.abs().max().item()), f"fused vs fp32 reference: {err}"
    assert (y0 + r - y).abs().max().item() <= 1e-5 * max(1.0, y.abs().max().item())
    if ys is not None:
        assert (ys - y).abs().max().item() <= 2e-5 * max(1.0, y.abs().max().item())
