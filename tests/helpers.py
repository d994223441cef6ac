"""Shared helpers for the parity tests (golden replay, model-dir cache)."""
import json
import tempfile
from functools import lru_cache
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
_TMP = tempfile.TemporaryDirectory(prefix="scb200_models_")


@lru_cache(maxsize=None)
def model_dir(arch: str, seed: int = 0, sharpen: float = 1.0) -> str:
    from speechcatcher_b200.synthetic import make_model_dir
    p = Path(_TMP.name) / f"{arch}_s{seed}_x{sharpen}"
    if not p.exists():
        make_model_dir(p, arch, seed=seed, sharpen=sharpen)
    return str(p)


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    calls = []
    for i in range(len(meta["calls"])):
        j = json.loads(str(z[f"c{i}_json"]))
        calls.append(dict(
            feats=z[f"c{i}_feats"] if f"c{i}_feats" in z else None,
            enc=z[f"c{i}_enc"] if f"c{i}_enc" in z else None,
            score=z[f"c{i}_score"], process_idx=int(z[f"c{i}_process_idx"]),
            yseq=j["yseq"], xpos=j["xpos"], results=j["results"]))
    trace = []
    j = 0
    while f"t{j}_dec" in z:
        trace.append(dict(dec=z[f"t{j}_dec"], ctc=z[f"t{j}_ctc"], comb=z[f"t{j}_comb"], T=int(z[f"t{j}_T"])))
        j += 1
    return meta, calls, trace


GOLDEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))
