#!/usr/bin/env python
"""How does the tcgen05 kind::f16 data path accumulate?  (GPU only; prints JSON.)

Feeds the existing bf16 tensor-core GEMM (sc_linear_bf16) with bf16-exact operands and compares the fp32 result with
(a) the exact fp64 dot product and (b) a sequential fp32 round-to-nearest accumulation.  A mean signed error well away
from zero on all-positive data means the accumulator truncates (round-toward-zero); the RMS tells how many bits survive.
This decides how the split-precision (fp16 hi/lo) GEMM of the precise mode has to accumulate.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechcatcher_b200 import _lib  # noqa: E402


def run(M, N, K, positive, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.rand(M, K, device="cuda", generator=g) if positive else torch.randn(M, K, device="cuda", generator=g)
    w = torch.rand(N, K, device="cuda", generator=g) if positive else torch.randn(N, K, device="cuda", generator=g)
    a16, w16 = a.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous()
    y = torch.empty(M, N, device="cuda")
    lib = _lib.load()
    _lib.check(lib.sc_linear_bf16(C.c_void_p(a16.data_ptr()), C.c_void_p(w16.data_ptr()), None, None,
                                  C.c_void_p(y.data_ptr()), None, M, N, K, 0, None), "linear_bf16")
    torch.cuda.synchronize()
    exact = a16.double() @ w16.double().t()
    # sequential fp32 RN accumulation in k order (what a plain FMA loop would give, products exact in fp32: bf16 x bf16)
    acc = torch.zeros(M, N, device="cuda")
    af, wf = a16.float(), w16.float()
    for k in range(K):
        acc = acc + af[:, k:k + 1] * wf[:, k].unsqueeze(0)
    rel = ((y.double() - exact) / exact.abs().clamp_min(1e-30))
    rel_seq = ((acc.double() - exact) / exact.abs().clamp_min(1e-30))
    scale = exact.abs().mean().item()
    return {"M": M, "N": N, "K": K, "positive": positive,
            "tc_mean_signed_rel": rel.mean().item(), "tc_rms_rel": rel.pow(2).mean().sqrt().item(),
            "tc_max_abs_err_over_scale": ((y.double() - exact).abs().max().item() / scale),
            "fp32seq_mean_signed_rel": rel_seq.mean().item(), "fp32seq_rms_rel": rel_seq.pow(2).mean().sqrt().item(),
            "ulp": 2.0 ** -24}


if __name__ == "__main__":
    out = []
    for K in (256, 2048):
        for positive in (True, False):
            out.append(run(256, 256, K, positive, 1))
    print(json.dumps(out, indent=1))
