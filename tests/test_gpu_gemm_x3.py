"""GPU: the split-precision tensor-core GEMM (sc_linear_x3: fp32 in, fp32 out, fp16 hi/lo operands on tcgen05) against
an fp64 PyTorch reference of the same op, next to the CUDA-core fp32 GEMM it replaces.

Bar: fp32-class accuracy -- the error against fp64 must stay within a small multiple of what the true-fp32 FMA kernel
(sc_linear_f32) shows on the same operands, and far below anything a bf16 operand rounding (2^-9) would give."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # (M, N, K, relu, residual)
    (128, 128, 64, 0, False),
    (128, 64, 256, 0, False),
    (333, 768, 256, 0, False),        # ragged M, QKV shape
    (77, 256, 2048, 0, True),         # FFN2 with in-place residual, small M (BN = 64 path)
    (130, 2048, 256, 1, False),       # FFN1 + ReLU
    (5, 1024, 256, 0, False),         # output layer, tiny M
    (64, 256, 4864, 0, False),        # embed.out
    (2560, 256, 256, 0, True),        # decoder projections with residual
    (2560, 768, 256, 0, False),
    (10752, 2048, 256, 1, False),     # encoder FFN1 at 256 streams x 1 block (BN = 128 path)
    (10752, 256, 2048, 0, True),
    (21504, 768, 256, 0, False),      # BN = 128, many tiles
    (3000, 256, 2304, 1, False),      # conv2 shape (dense A)
    (640, 512, 256, 0, False),        # cross K|V
]


def _planes(w):
    from speechcatcher_b200.weights import split_f16
    return split_f16(w)


@pytest.mark.parametrize("M,N,K,relu,res", SHAPES)
def test_linear_x3_fp32_class_accuracy(M, N, K, relu, res):
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g, device="cuda") * 1.7 + 0.1
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    r = torch.randn(M, N, generator=g, device="cuda")
    w2 = _planes(w)
    y = r.clone() if res else torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.sc_linear_x3(a.data_ptr(), w2.data_ptr(), bias.data_ptr(), y.data_ptr() if res else None,
                                y.data_ptr(), M, N, K, relu, None), "linear_x3")
    y32 = r.clone() if res else torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.sc_linear_f32(a.data_ptr(), w.data_ptr(), bias.data_ptr(), y32.data_ptr() if res else None,
                                 y32.data_ptr(), M, N, K, relu, None), "linear_f32")
    torch.cuda.synchronize()
    want = a.double() @ w.double().t() + bias.double()
    if relu:
        want = want.relu()
    if res:
        want = want + r.double()
    assert torch.isfinite(y).all()
    err = (y.double() - want).abs()
    err32 = (y32.double() - want).abs()
    scale = want.abs().max().item()
    # fp32-class: max error within 4x of the CUDA-core fp32 kernel's (plus one fp32 ulp of the row scale), rms likewise
    assert err.max().item() <= 4.0 * err32.max().item() + 2.0 ** -22 * scale, \
        f"x3 max err {err.max().item():.3e} vs fp32 kernel {err32.max().item():.3e} (scale {scale:.2f})"
    assert err.pow(2).mean().sqrt().item() <= 4.0 * err32.pow(2).mean().sqrt().item() + 2.0 ** -24 * scale
    # and nowhere near a bf16 / fp16 single-pass result
    assert err.max().item() <= 2e-5 * max(1.0, scale)


def test_linear_x3_small_and_large_magnitudes():
    """Operands far from 1: tiny values (below the fp16 normal range) and large ones (up to ~1e3) keep fp32-class
    *absolute* accuracy relative to the output scale."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 256, 256, 256
    for a_scale, w_scale in [(1e-4, 1.0), (300.0, 0.05), (1.0, 1e-5), (1e-6, 1e-3)]:
        a = torch.randn(M, K, generator=g, device="cuda") * a_scale
        w = torch.randn(N, K, generator=g, device="cuda") * w_scale
        y = torch.empty(M, N, device="cuda")
        _lib.check(lib.sc_linear_x3(a.data_ptr(), _planes(w).data_ptr(), None, None, y.data_ptr(), M, N, K, 0, None), "x3")
        torch.cuda.synchronize()
        want = a.double() @ w.double().t()
        scale = want.abs().max().item()
        # operands below the fp16 normal range (2^-14) keep an ABSOLUTE precision of ~2^-36 (hi is a subnormal fp16,
        # lo = (x - hi) * 2^11 sits at the bottom of the normal range) instead of a relative one
        floor = 4.0 * K * 2.0 ** -36 * (a.abs().max().item() + w.abs().max().item())
        assert (y.double() - want).abs().max().item() <= 3e-6 * scale + floor, (a_scale, w_scale)


def _act_planes(x):
    """Split fp16 planes of an activation matrix, with the device's arithmetic (x3_split.cuh)."""
    hi = x.to(torch.float16)
    lo = ((x - hi.float()) * 2048.0).to(torch.float16)
    return torch.stack([hi, lo]).contiguous()


@pytest.mark.parametrize("M,N,K,relu,res", [(333, 768, 256, 0, False), (2560, 256, 2048, 0, True), (10752, 2048, 256, 1, False),
                                            (77, 1024, 256, 0, False), (3000, 256, 4864, 0, False)])
def test_linear_x3_planes_in_equals_converter_path(M, N, K, relu, res):
    """The TMA-fed form (activation already stored as split planes by its producer) computes exactly what the converter
    form computes from the fp32 rows: same split, same tensor-core instruction sequence."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    cap = M + 200                                             # row capacity of the plane buffer > M (stale rows behind)
    a = torch.randn(cap, K, generator=g, device="cuda") * 1.3
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    r = torch.randn(M, N, generator=g, device="cuda")
    w2, a2 = _planes(w), _act_planes(a)
    y0 = r.clone() if res else torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.sc_linear_x3(a.data_ptr(), w2.data_ptr(), bias.data_ptr(), y0.data_ptr() if res else None, y0.data_ptr(),
                                M, N, K, relu, None), "x3")
    y1 = r.clone() if res else torch.full((M, N), float("nan"), device="cuda")
    yp = torch.zeros(2, M, N, dtype=torch.float16, device="cuda")
    _lib.check(lib.sc_linear_x3_planes(a2.data_ptr(), cap * K, cap, w2.data_ptr(), bias.data_ptr(),
                                       y1.data_ptr() if res else None, y1.data_ptr(), yp.data_ptr(), M * N,
                                       M, N, K, relu, 1, None), "x3_planes")
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)
    # the plane output represents the fp32 result to 2^-22 relative (2^-36 absolute for tiny values)
    rec = yp[0].double() + yp[1].double() / 2048.0
    assert ((rec - y1.double()).abs() <= 2.0 ** -21 * y1.double().abs() + 2.0 ** -34).all()
    # planes only (no fp32 output), as FFN1 hands its result to FFN2
    if not res:
        yp2 = torch.zeros_like(yp)
        _lib.check(lib.sc_linear_x3_planes(a2.data_ptr(), cap * K, cap, w2.data_ptr(), bias.data_ptr(), None, None,
                                           yp2.data_ptr(), M * N, M, N, K, relu, 1, None), "x3_planes_only")
        torch.cuda.synchronize()
        assert torch.equal(yp, yp2)


PERSISTENT_SHAPES = [
    # (M, N, K, relu, out) -- out: "f32" store, "res" in-place residual (fp32 add at the L2), "planes"
    (10752, 768, 256, 0, "f32"),      # encoder QKV: 84 x 6 tiles, A-resident, several tiles per CTA
    (8610, 2048, 256, 1, "planes"),   # encoder FFN1 (ragged M: 205 blocks x 42 rows), planes out
    (8610, 256, 256, 0, "res"),       # encoder O-projection, residual in place
    (8610, 256, 2048, 0, "res"),      # encoder FFN2: streamed A, 16 accumulator chunks
    (126, 768, 256, 0, "f32"),        # three blocks of one stream: one row tile
    (42, 2048, 256, 1, "planes"),
    (42, 256, 2048, 0, "res"),
    (1000, 128, 384, 0, "f32"),       # odd chunk count (3), one column tile
    (40000, 256, 256, 0, "res"),      # more row tiles than SMs: a CTA walks several row tiles (A refill behind the MMAs)
    (20000, 512, 256, 1, "planes"),
]


@pytest.mark.parametrize("M,N,K,relu,out", PERSISTENT_SHAPES)
def test_linear_x3_persistent_kernel(M, N, K, relu, out):
    """The persistent kernel (kernels_gemm_x3p.cu) against an fp64 reference and against the per-tile kernel on the same
    plane operands: same products, a different (chunked, round-to-nearest) accumulation order, so the two agree to
    fp32 rounding level, and both carry fp32-class error."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    a = torch.randn(M, K, generator=g, device="cuda") * 1.3 + 0.05
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    r = torch.randn(M, N, generator=g, device="cuda")
    w2, a2 = _planes(w), _act_planes(a)
    a_rec = a2[0].double() + a2[1].double() / 2048.0          # what the planes represent
    w_rec = w2.view(2, N, K)[0].double() + w2.view(2, N, K)[1].double() / 2048.0
    want = a_rec @ w_rec.t() + bias.double()
    if relu:
        want = want.relu()
    if out == "res":
        want = want + r.double()
    outs = []
    for kernel in (1, 2, 3):
        y = r.clone() if out == "res" else torch.full((M, N), float("nan"), device="cuda")
        yp = torch.full((2, M + 3, N), 7.0, dtype=torch.float16, device="cuda")     # 3 guard rows behind the last one
        planes = out == "planes"
        _lib.check(lib.sc_linear_x3_planes(a2.data_ptr(), M * K, M, w2.data_ptr(), bias.data_ptr(),
                                           y.data_ptr() if out == "res" else None, None if planes else y.data_ptr(),
                                           yp.data_ptr() if planes else None, (M + 3) * N, M, N, K, relu, kernel, None),
                   f"x3_planes kernel {kernel}")
        torch.cuda.synchronize()
        if planes:
            assert (yp[:, M:] == 7.0).all()                   # rows beyond M are clipped
            y = yp[0, :M].float() + yp[1, :M].float() / 2048.0
        assert torch.isfinite(y).all()
        outs.append(y.double())
    scale = want.abs().max().item()
    err_tile = (outs[0] - want).abs()
    tol_planes = 2.0 ** -21 * scale if out == "planes" else 0.0
    for which, y in (("persistent", outs[1]), ("persistent, smem operands", outs[2])):
        err_pers = (y - want).abs()
        assert err_pers.max().item() <= 2.0 * err_tile.max().item() + 2.0 ** -22 * scale + tol_planes, \
            f"{which}: max err {err_pers.max().item():.3e} vs per-tile {err_tile.max().item():.3e} (scale {scale:.2f})"
        assert err_pers.pow(2).mean().sqrt().item() <= 2.0 * err_tile.pow(2).mean().sqrt().item() + 2.0 ** -24 * scale
        assert (outs[0] - y).abs().max().item() <= 2.0 ** -20 * scale
    # kernel 2 (engine's choice of persistent form; SCB_X3T=1 selects A in tensor memory for K = 256) and kernel 3 (operands
    # in shared memory) run the same instruction sequence on the same planes
    assert torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("M,N,relu,out,n_rows", [(8610, 768, 0, "f32", None), (8610, 2048, 1, "planes", None),
                                                 (2560, 768, 0, "f32", 2480), (2560, 256, 0, "f32", 130),
                                                 (2560, 2048, 1, "planes", 2333), (2560, 1024, 0, "f32", 1),
                                                 (42, 768, 0, "f32", None), (40000, 256, 0, "f32", None)])
def test_linear_x3_layernorm_prologue(M, N, relu, out, n_rows):
    """LayerNorm in the prologue of the persistent GEMM == sc_layernorm_split followed by the persistent GEMM on the
    planes (bit for bit: same LayerNorm arithmetic, same split, same instruction sequence); with a device-side row count
    only rows below it are defined."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    K = 256
    g = torch.Generator(device="cuda").manual_seed(M + N + (n_rows or 0))
    x = torch.randn(M, K, generator=g, device="cuda") * 2.5 + 0.3
    lw = 1.0 + 0.1 * torch.randn(K, generator=g, device="cuda")
    lb = 0.1 * torch.randn(K, generator=g, device="cuda")
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    w2 = _planes(w)
    nr = torch.tensor([n_rows], dtype=torch.int32, device="cuda") if n_rows is not None else None
    planes = out == "planes"
    xp = torch.zeros(2, M, K, dtype=torch.float16, device="cuda")
    _lib.check(lib.sc_layernorm_split(x.data_ptr(), lw.data_ptr(), lb.data_ptr(), xp.data_ptr(), M * K, M, K, None), "ln_split")
    y0 = torch.full((M, N), float("nan"), device="cuda")
    yp0 = torch.zeros(2, M, N, dtype=torch.float16, device="cuda")
    _lib.check(lib.sc_linear_x3_planes(xp.data_ptr(), M * K, M, w2.data_ptr(), bias.data_ptr(), None,
                                       None if planes else y0.data_ptr(), yp0.data_ptr() if planes else None, M * N,
                                       M, N, K, relu, 2, None), "x3_planes persistent")
    y1 = torch.full((M, N), float("nan"), device="cuda")
    yp1 = torch.zeros(2, M, N, dtype=torch.float16, device="cuda")
    _lib.check(lib.sc_linear_x3_ln(x.data_ptr(), lw.data_ptr(), lb.data_ptr(), w2.data_ptr(), bias.data_ptr(),
                                   None if planes else y1.data_ptr(), yp1.data_ptr() if planes else None, M * N,
                                   M, N, relu, nr.data_ptr() if nr is not None else None, None), "x3_ln")
    torch.cuda.synchronize()
    rows = M if n_rows is None else n_rows
    if planes:
        assert torch.equal(yp0[:, :rows], yp1[:, :rows])
    else:
        assert torch.isfinite(y1[:rows]).all()
        assert torch.equal(y0[:rows], y1[:rows])
    # against fp64
    xn = torch.nn.functional.layer_norm(x.double(), (K,), lw.double(), lb.double(), eps=1e-12)
    want = xn @ w.double().t() + bias.double()
    if relu:
        want = want.relu()
    got = (yp1[0].double() + yp1[1].double() / 2048.0) if planes else y1.double()
    assert (got[:rows] - want[:rows]).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_layernorm_split_planes():
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    rows = 1234
    x = torch.randn(rows, 256, generator=g, device="cuda") * 3 + 0.5
    w = 1.0 + 0.1 * torch.randn(256, generator=g, device="cuda")
    b = 0.1 * torch.randn(256, generator=g, device="cuda")
    y = torch.empty_like(x)
    _lib.check(lib.sc_layernorm_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, 256, None), "ln")
    yp = torch.zeros(2, rows, 256, dtype=torch.float16, device="cuda")
    _lib.check(lib.sc_layernorm_split(x.data_ptr(), w.data_ptr(), b.data_ptr(), yp.data_ptr(), rows * 256, rows, 256, None), "ln_split")
    torch.cuda.synchronize()
    assert torch.equal(yp, _act_planes(y))          # the fp32 LayerNorm's result, split
