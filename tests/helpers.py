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
def model_dir(arch: str, seed: int = 0, sharpen: float = 1.0, eos_bias: float = 0.0) -> str:
    from speechcatcher_b200.synthetic import make_model_dir
    p = Path(_TMP.name) / f"{arch}_s{seed}_x{sharpen}_e{eos_bias}"
    if not p.exists():
        make_model_dir(p, arch, seed=seed, sharpen=sharpen, eos_bias=eos_bias)
    return str(p)


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    calls = []
    for i in range(len(meta["calls"])):
        if meta.get("lite"):
            # long-utterance goldens: beams only (int16 arrays), no feature / encoder tensors
            calls.append(dict(
                feats=None, enc=None, called=bool(z[f"c{i}_called"]), n_feat=int(z[f"c{i}_nfeat"]), n_enc=int(z[f"c{i}_nenc"]),
                score=z[f"c{i}_score"], process_idx=int(z[f"c{i}_process_idx"]),
                yseq=z[f"c{i}_yseq"].astype(int).tolist(), xpos=z[f"c{i}_xpos"].astype(int).tolist(),
                results=json.loads(str(z[f"c{i}_results"]))))
            continue
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


class OracleGroup:
    """StreamGroup protocol (push / last_plan / results / reset) over N independent CPU oracles -- the checker the
    file-level tests run `speechcatcher_b200.recognize` against.  Result tuples are built with plain loops that follow
    speech2text_streaming.py:466-537 line by line (not the product's vectorised assembly)."""

    def __init__(self, model_dir_, n_streams, beam_size=5, max_seconds=1e9, max_chunk=1 << 30):
        from oracle.speech2text import OracleSpeech2Text
        self.n_streams, self.max_seconds, self.max_chunk = n_streams, max_seconds, max_chunk
        self.o = [OracleSpeech2Text(model_dir_, beam_size=beam_size) for _ in range(n_streams)]
        self.called = [False] * n_streams
        self.token_list = None

    def reset(self, streams=None):
        for s in (range(self.n_streams) if streams is None else streams):
            self.o[s].reset()

    def push(self, ids, chunks, is_final):
        for s, c, f in zip(ids, chunks, is_final):
            self.o[s](np.asarray(c, np.float32), is_final=bool(f), finalize_all=False)
            self.called[s] = self.o[s].last_feats is not None

    def last_plan(self, s):
        from types import SimpleNamespace
        return SimpleNamespace(called=int(self.called[s]))

    def results(self, s, is_final, finalize_all, token_list=None):
        hyps = self.o[s].hyps
        if not is_final or not finalize_all:
            hyps = [h for h in hyps if h.yseq[-1] == 1023]
        out = []
        for h in hyps:
            ids, pos = list(h.yseq[1:]), list(h.xpos[1:])
            if not is_final:
                ids, pos = [], []
            elif ids and ids[-1] == 1023:
                ids, pos = ids[:-1], pos[:-1]
            keep = [i for i, t in enumerate(ids) if t not in (0, 1, 1023)]
            ids, pos = [ids[i] for i in keep], [pos[i] for i in keep]
            toks = [str(t) for t in ids]
            out.append((" ".join(toks), toks, ids, pos, dict(yseq=list(h.yseq), score=h.score, xpos=list(h.xpos))))
        return out
