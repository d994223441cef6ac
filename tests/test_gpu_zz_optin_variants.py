"""GPU: the measured-and-rejected kernel variants that stay in the tree as opt-in experiments must
keep reproducing the reference: each one replays two goldens (exact n-best, timestamps, order) in a fresh process with
its environment switch set (the switches are read once per process)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

REPO = Path(__file__).resolve().parent.parent

VARIANTS = {
    "decode_chain_kernel": {"SCB_X3_CHAIN": "1"},                 # kernels_chain_x3.cu
    "gemm_a_in_tensor_memory": {"SCB_X3T": "1"},                  # kernels_gemm_x3t.cu
    "layernorm_prologue": {"SCB_X3_LN_FUSED": "1"},               # kernels_gemm_x3p.cu <A_RES, LN>
    "attention_warp_per_head": {"SCB_ATTN_WARP_HEAD": "1"},       # kernels_attn_x3.cu <WH>
    "decode_projections_two_ctas_per_sm": {"SCB_X3_DEC2": "1"},   # kernels_gemm_x3.cu <64, 2, 2, true, 2>
    "decode_projections_persistent": {"SCB_X3_DEC_PERSIST": "2"},
    "per_tile_gemm_only": {"SCB_X3_PERSIST_MIN_M": "0"},
    "direct_store_epilogue": {"SCB_XP_DIRECT": "1"},
    "cuda_graph_replay": {"SCB_GRAPH": "3", "SCB_OWN_STREAM": "1"},
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_replays_goldens(name):
    env = dict(os.environ)
    env.update(VARIANTS[name])
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-m", "gpu", "-x",
                        "-k", "xl_d4_b10_cli or m_d2_b5_6s or xl_b10_4s"], cwd=REPO, env=env, capture_output=True, text=True)
    assert r.returncode == 0, f"{name} {VARIANTS[name]}:\n{r.stdout[-3000:]}\n{r.stderr[-1500:]}"
