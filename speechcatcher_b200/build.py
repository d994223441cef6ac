"""Build the CUDA extension in-tree: speechcatcher_b200/libscb200.so (sm_100a only).

`python -m speechcatcher_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libscb200.so"
SOURCES = ["kernels_gemm.cu", "kernels_gemm_tc.cu", "kernels_gemm_x3.cu", "kernels_gemm_x3p.cu", "kernels_gemm_x3t.cu", "kernels_chain_x3.cu", "kernels_frontend.cu", "kernels_encoder.cu",
           "kernels_search.cu", "kernels_attn_mma.cu", "kernels_attn_f32.cu", "kernels_attn_x3.cu", "kernels_ffn_fused.cu", "segmenter.cu", "engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [HERE.parent / "include" / "speechcatcher_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = objdir / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("SCB_NVCC_EXTRA", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(OUT), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-lcudart"]          # the driver API is resolved at run time (cudaGetDriverEntryPoint): no libcuda dependency
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
