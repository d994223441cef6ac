"""Micro-benchmark of the tcgen05 GEMM (sc_linear_bf16) on the encoder / decoder shapes of the XL model.

Not a pytest module: run on a B200 as `python scripts/gemm_bench.py`.  Each shape rotates over enough distinct
activation buffers to exceed L2 only when `--cold` is given; the default (warm) mirrors the engine, where the A
operand has just been written by the preceding kernel.  Times are CUDA-event averages over `--iters` launches.
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechcatcher_b200 import _lib  # noqa: E402

SHAPES = [  # (name, M, N, K, relu, residual, f32_out, bf16_out)
    ("enc_qkv", 10752, 768, 256, 0, 0, 0, 1),
    ("enc_o", 10752, 256, 256, 0, 1, 1, 0),
    ("enc_ffn1", 10752, 2048, 256, 1, 0, 0, 1),
    ("enc_ffn2", 10752, 256, 2048, 0, 1, 1, 0),
    ("enc_qkv_s", 2688, 768, 256, 0, 0, 0, 1),
    ("enc_o_s", 2688, 256, 256, 0, 1, 1, 0),
    ("enc_ffn1_s", 2688, 2048, 256, 1, 0, 0, 1),
    ("enc_ffn2_s", 2688, 256, 2048, 0, 1, 1, 0),
    ("dec_qkv", 2560, 768, 256, 0, 0, 0, 1),
    ("dec_o", 2560, 256, 256, 0, 1, 1, 0),
    ("dec_ffn1", 2560, 2048, 256, 1, 0, 0, 1),
    ("dec_ffn2", 2560, 256, 2048, 0, 1, 1, 0),
    ("dec_ffn1_s", 640, 2048, 256, 1, 0, 0, 1),
    ("dec_ffn2_s", 640, 256, 2048, 0, 1, 1, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cold", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream(dev)
    out = {}
    for name, M, N, K, relu, res, f32o, b16o in SHAPES:
        per = M * K * 2 + (M * N * 4 if f32o else 0) + (M * N * 2 if b16o else 0) + (M * N * 4 if res else 0)
        nbuf = max(2, (400 << 20) // per) if args.cold else 2
        g = torch.Generator(device=dev).manual_seed(1)
        xs = [torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16) for _ in range(nbuf)]
        w = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
        b = torch.randn(N, device=dev, generator=g)
        rs = [torch.randn(M, N, device=dev, generator=g) for _ in range(nbuf)] if res else None
        ys = [torch.empty(M, N, device=dev) for _ in range(nbuf)] if f32o else None
        yb = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)] if b16o else None

        def launch(i):
            j = i % nbuf
            rc = lib.sc_linear_bf16(C.c_void_p(xs[j].data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                    C.c_void_p(rs[j].data_ptr()) if res else None,
                                    C.c_void_p(ys[j].data_ptr()) if f32o else None,
                                    C.c_void_p(yb[j].data_ptr()) if b16o else None, M, N, K, relu,
                                    C.c_void_p(st.cuda_stream))
            _lib.check(rc, "sc_linear_bf16")

        for i in range(10):
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        us = 1000.0 * e0.elapsed_time(e1) / args.iters
        # correctness spot check against torch on the last buffer used
        j = (args.iters - 1) % nbuf
        ref = xs[j].float() @ w.float().t() + b
        if relu:
            ref = ref.relu()
        if res:
            ref = ref + rs[j]
        got = ys[j] if f32o else yb[j].float()
        err = float((got - ref).abs().max())
        out[name] = {"M": M, "N": N, "K": K, "us": round(us, 2), "tflops": round(2.0 * M * N * K / us / 1e6, 1),
                     "gbs": round(per / us / 1e3, 1), "max_err": err}
        print(name, out[name], flush=True)
    # fused FFN (FFN1 -> ReLU -> FFN2 in one kernel) against the sum of the two GEMMs above
    for name, M, splits in (("enc_ffn_fused", 10752, 1), ("enc_ffn_fused_s", 2688, 1), ("enc_ffn_fused_s_x4", 2688, 4),
                            ("dec_ffn_fused", 2560, 1), ("dec_ffn_fused_x4", 2560, 4), ("dec_ffn_fused_s_x8", 640, 8)):
        F, D = 2048, 256
        g = torch.Generator(device=dev).manual_seed(2)
        xs = [torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(2)]
        w1 = (torch.randn(F, D, device=dev, generator=g) * 0.05).to(torch.bfloat16)
        w2 = (torch.randn(D, F, device=dev, generator=g) * 0.02).to(torch.bfloat16)
        b1 = torch.randn(F, device=dev, generator=g)
        b2 = torch.randn(D, device=dev, generator=g)
        ys = [torch.randn(M, D, device=dev, generator=g) for _ in range(2)]

        def launch(i):
            j = i % 2
            _lib.check(lib.sc_ffn_bf16(C.c_void_p(xs[j].data_ptr()), C.c_void_p(w1.data_ptr()), C.c_void_p(b1.data_ptr()),
                                       C.c_void_p(w2.data_ptr()), C.c_void_p(b2.data_ptr()), C.c_void_p(ys[j].data_ptr()),
                                       1, M, F, splits, C.c_void_p(st.cuda_stream)), "sc_ffn_bf16")

        for i in range(10):
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        us = 1000.0 * e0.elapsed_time(e1) / args.iters
        out[name] = {"M": M, "us": round(us, 2), "tflops": round(4.0 * M * F * D / us / 1e6, 1)}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
