"""Device-side timeline of the fused FFN kernel (first CTA): where a hidden chunk's time goes.
Run on a B200: python scripts/ffn_timeline.py [M]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechcatcher_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2688
F, D = 2048, 256
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16)
w1 = (torch.randn(F, D, device=dev, generator=g) * 0.05).to(torch.bfloat16)
w2 = (torch.randn(D, F, device=dev, generator=g) * 0.02).to(torch.bfloat16)
b1 = torch.randn(F, device=dev, generator=g)
b2 = torch.randn(D, device=dev, generator=g)
y = torch.randn(M, D, device=dev, generator=g)
st = torch.zeros(1024, dtype=torch.int64, device=dev)
for _ in range(3):
    _lib.check(lib.sc_ffn_bf16_timeline(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), y.data_ptr(),
                                        1, M, F, st.data_ptr(), None), "timeline")
torch.cuda.synchronize()
t = st.cpu().tolist()
t0 = t[0]
print(f"A tile loaded: +{t[1] - t0} clk; acc2 complete: +{t[2] - t0}; final epilogue done: +{t[3] - t0}")
print(" j | MMA: wait-start  H_j-ready  G2-issued  G1(j+2)-issued | EPI: wait-start acc1-done  H-free   E1-done")
for j in range(F // 128):
    m = [t[8 + j * 4 + i] - t0 for i in (0, 3, 1, 2)]
    e = [t[128 + j * 4 + i] - t0 for i in range(4)]
    print(f"{j:2d} | {m[0]:10d} {m[1]:10d} {m[2]:10d} {m[3]:10d}       | {e[0]:10d} {e[1]:10d} {e[2]:10d} {e[3]:10d}")
print("stage | load issued | MMA thread: starts waiting, data there | issue->there")
for c in list(range(0, 12)) + list(range(32, 44)):
    a, b, d = t[256 + c] - t0, t[384 + c] - t0, t[512 + c] - t0
    print(f"{c:4d} | {a:10d} | {b:10d} {d:10d} | {d - a:6d}")
