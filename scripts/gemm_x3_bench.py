#!/usr/bin/env python
"""Isolated timing of the precise-mode GEMMs (GPU only): sc_linear_x3_planes (TMA-fed, split planes), sc_linear_x3
(converter path), sc_linear_f32 (CUDA cores) and sc_linear_bf16 on the shapes of the hot path.  CUDA events, 20 launches
after 3 warm-ups, a 256 MB L2 flush between launches.  Prints one JSON line per shape."""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechcatcher_b200 import _lib  # noqa: E402
from speechcatcher_b200.weights import split_f16  # noqa: E402

SHAPES = [  # name, M, N, K, relu, residual
    ("enc_qkv", 10752, 768, 256, 0, 0), ("enc_o", 10752, 256, 256, 0, 1), ("enc_ffn1", 10752, 2048, 256, 1, 0),
    ("enc_ffn2", 10752, 256, 2048, 0, 1), ("enc_qkv_2blk", 21504, 768, 256, 0, 0), ("enc_ffn1_2blk", 21504, 2048, 256, 1, 0),
    ("enc_ffn2_2blk", 21504, 256, 2048, 0, 1),
    ("dec_qkv", 2560, 768, 256, 0, 0), ("dec_o", 2560, 256, 256, 0, 1), ("dec_ffn1", 2560, 2048, 256, 1, 0),
    ("dec_ffn2", 2560, 256, 2048, 0, 1), ("dec_out", 2560, 1024, 256, 0, 0),
    ("embed_out", 2816, 256, 4864, 0, 0), ("cross_kv", 4096, 512, 256, 0, 0),
]


def planes(x):
    hi = x.to(torch.float16)
    return torch.stack([hi, ((x - hi.float()) * 2048.0).to(torch.float16)]).contiguous()


def timeit(fn, flush, n=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2] * 1e3


def main():
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, M, N, K, relu, res in SHAPES:
        a = torch.randn(M, K, generator=g, device="cuda")
        w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
        bias = torch.randn(N, generator=g, device="cuda")
        y = torch.zeros(M, N, device="cuda")
        yp = torch.zeros(2, M, N, dtype=torch.float16, device="cuda")
        a2, w2 = planes(a), split_f16(w)
        a16, w16 = a.to(torch.bfloat16), w.to(torch.bfloat16)
        r = y.data_ptr() if res else None
        out_planes = relu and not res          # FFN1-like: the result is handed over as planes only
        t = {}
        t["x3_planes"] = timeit(lambda: _lib.check(lib.sc_linear_x3_planes(
            a2.data_ptr(), M * K, M, w2.data_ptr(), bias.data_ptr(), r, None if out_planes else y.data_ptr(),
            yp.data_ptr() if out_planes else None, M * N, M, N, K, relu, 1, None)), flush)
        if N % 128 == 0 and K % 128 == 0 and K >= 256:
            t["x3_persistent"] = timeit(lambda: _lib.check(lib.sc_linear_x3_planes(
                a2.data_ptr(), M * K, M, w2.data_ptr(), bias.data_ptr(), r, None if out_planes else y.data_ptr(),
                yp.data_ptr() if out_planes else None, M * N, M, N, K, relu, 2, None)), flush)
        t["x3_convert"] = timeit(lambda: _lib.check(lib.sc_linear_x3(a.data_ptr(), w2.data_ptr(), bias.data_ptr(), r, y.data_ptr(),
                                                                     M, N, K, relu, None)), flush)
        t["f32_simt"] = timeit(lambda: _lib.check(lib.sc_linear_f32(a.data_ptr(), w.data_ptr(), bias.data_ptr(), r, y.data_ptr(),
                                                                    M, N, K, relu, None)), flush, n=5)
        t["bf16"] = timeit(lambda: _lib.check(lib.sc_linear_bf16(a16.data_ptr(), w16.data_ptr(), bias.data_ptr(), r, y.data_ptr(),
                                                                 None, M, N, K, relu, None)), flush)
        fl = 2.0 * M * N * K
        print(json.dumps({"shape": name, "M": M, "N": N, "K": K,
                          "us": {k: round(v, 1) for k, v in t.items()},
                          "tflops_algorithmic": {k: round(fl / v / 1e6, 1) for k, v in t.items()}}), flush=True)


if __name__ == "__main__":
    main()
