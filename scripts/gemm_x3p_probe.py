#!/usr/bin/env python
"""Timing probe of the persistent split-fp16 GEMM (kernels_gemm_x3p.cu) on the encoder shapes of one push (205 blocks x
42 rows), warm L2 (back-to-back launches, as inside the encoder stack) and cold (256 MB flush between launches).
SCB_XP_DBG=1 drops the output stores, =3 also the TMEM drain (timing experiments: what bounds a tile)."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechcatcher_b200 import _lib  # noqa: E402
from speechcatcher_b200.weights import split_f16  # noqa: E402

SHAPES = [("enc_qkv", 8610, 768, 256, 0, 0), ("enc_o", 8610, 256, 256, 0, 1), ("enc_ffn1", 8610, 2048, 256, 1, 0),
          ("enc_ffn2", 8610, 256, 2048, 0, 1), ("enc_ffn1_big", 21504, 2048, 256, 1, 0)]


def planes(x):
    hi = x.to(torch.float16)
    return torch.stack([hi, ((x - hi.float()) * 2048.0).to(torch.float16)]).contiguous()


def main():
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    kernels = [int(k) for k in os.environ.get("PROBE_KERNELS", "2,1").split(",")]
    for name, M, N, K, relu, res in SHAPES:
        a = torch.randn(M, K, generator=g, device="cuda")
        w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
        bias = torch.randn(N, generator=g, device="cuda")
        y = torch.zeros(M, N, device="cuda")
        yp = torch.zeros(2, M, N, dtype=torch.float16, device="cuda")
        a2, w2 = planes(a), split_f16(w)
        out_planes = bool(relu)
        out = {"shape": name, "M": M, "N": N, "K": K, "dbg": os.environ.get("SCB_XP_DBG", "0")}
        for kernel in kernels:
            def fn():
                _lib.check(lib.sc_linear_x3_planes(a2.data_ptr(), M * K, M, w2.data_ptr(), bias.data_ptr(),
                                                   y.data_ptr() if res else None, None if out_planes else y.data_ptr(),
                                                   yp.data_ptr() if out_planes else None, M * N, M, N, K, relu, kernel, None))
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            warm = e0.elapsed_time(e1) * 1e3 / 20
            cold = []
            for _ in range(8):
                flush.zero_()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(); fn(); c1.record()
                torch.cuda.synchronize()
                cold.append(c0.elapsed_time(c1) * 1e3)
            cold.sort()
            tag = "persistent" if kernel == 2 else "per_tile"
            out[tag] = {"warm_us": round(warm, 1), "cold_us": round(cold[len(cold) // 2], 1),
                        "tflops_algorithmic_warm": round(2.0 * M * N * K / warm / 1e6, 1)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
