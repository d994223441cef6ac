"""Generate tests/golden/feats_input.json by running the REFERENCE with pre-computed features (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.gen_golden_feats

The reference's own tests (tests/test_speech2text_streaming.py:93-190) drive `Speech2TextStreaming` with 2-D feature
chunks of 100 frames instead of waveforms (speech2text_streaming.py:438-450: normalise, skip the frontend).  This pins
that input mode: per call the beam (yseq / xpos / fp64 score / process_idx) and the returned ids, for 2-D chunks (with
MVN statistics), a short final chunk, a sub-3-frame chunk (encoder skipped) and the 3-D batched form (not normalised).
Feature chunks are regenerated from their seeds by the tests.
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle.speech2text import OracleSpeech2Text  # noqa: E402
from speechcatcher_b200.synthetic import make_model_dir  # noqa: E402

# name, arch, beam, eos_bias, seed, frames per call, final flags, batched (3-D) input
CASES = [
    ("feats2d_m_d2_100x3", "m_d2", 3, 0.0, 1, [100, 100, 100], [False, False, True], False),
    ("feats2d_xl_d4_ragged", "xl_d4", 5, 7.0, 2, [100, 37, 2, 64, 100, 9], [False, False, False, False, False, True], False),
    ("feats3d_m_d2", "m_d2", 5, 7.0, 3, [80, 80, 80, 50], [False, False, False, True], True),
    ("feats2d_m_d2_midfinal", "m_d2", 5, 7.0, 4, [100, 100, 60, 100, 100], [False, False, True, False, True], False),
]


def feature_chunks(seed, frames):
    """Feature-like values around the synthetic MVN statistics (mean ~ -8, std ~ 2) so normalised inputs are O(1)."""
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal((n, 80)) * 2.0 - 8.0).astype(np.float32) for n in frames]


def main():
    sys.path.insert(0, "/root/reference")
    from speechcatcher.speech2text_streaming import Speech2TextStreaming as Ref
    out = []
    for name, arch, beam, eos_bias, seed, frames, finals, batched in CASES:
        with tempfile.TemporaryDirectory() as td:
            md = make_model_dir(td, arch, seed=0, eos_bias=eos_bias)
            ref, orc = Ref(md, beam_size=beam, device="cpu"), OracleSpeech2Text(md, beam_size=beam)
            calls = []
            for f, fin in zip(feature_chunks(seed, frames), finals):
                x = f[None] if batched else f
                r = ref(x.copy(), is_final=fin, finalize_all=fin)
                o = orc(x.copy(), is_final=fin, finalize_all=fin)
                hy = ref.beam_state.hypotheses
                rec = dict(yseq=[h.yseq.tolist() for h in hy], xpos=[h.xpos.tolist() for h in hy],
                           score=[float(h.score) for h in hy], process_idx=int(ref.beam_search.process_idx),
                           results=[[int(t) for t in x_[2]] for x_ in r])
                assert rec["yseq"] == [list(h.yseq) for h in orc.hyps], name
                assert rec["xpos"] == [list(h.xpos) for h in orc.hyps], name
                assert np.allclose(rec["score"], [h.score for h in orc.hyps], atol=1e-3), name
                assert rec["results"] == [list(x_[2]) for x_ in o] and rec["process_idx"] == orc.search.process_idx, name
                calls.append(rec)
        out.append(dict(name=name, arch=arch, beam=beam, eos_bias=eos_bias, seed=seed, frames=frames, finals=finals,
                        batched=batched, calls=calls))
        print(name, "final hypothesis length", len(calls[-1]["yseq"][0]), "results", len(calls[-1]["results"]))
    p = REPO / "tests" / "golden" / "feats_input.json"
    p.write_text(json.dumps(out))
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
