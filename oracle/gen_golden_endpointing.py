"""Generate tests/golden/endpointing.json by running the REFERENCE's own segmentation code (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage (needs /root/reference):

    python -m oracle.gen_golden_endpointing

`speechcatcher/simple_endpointing.py` imports `ffmpeg` and `python_speech_features` at module level; both are absent
from this image, so they are stubbed in sys.modules.  What the goldens pin:
  * `search_*`  : the reference's BeamSearch.search (simple_endpointing.py:43-69) on seeded energy curves,
  * `core_*`    : the reference's segment_speech (:72-137: scipy gaussian_filter1d, search, 180 s cap) on seeded int16
                  audio with `logfbank` supplied by oracle.endpointing.psf_logfbank -- the feature step itself stays
                  "parity unpinned" (python_speech_features is not installed anywhere we can run).
The curves/audio are regenerated from their seeds by the tests (only seeds, parameters and cuts are stored).
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle.endpointing import (CutSearchOracle, psf_logfbank, segment_speech_oracle)  # noqa: E402


def energy_curve(seed: int, n: int, shift: float = 0.0) -> np.ndarray:
    """A smoothed-energy-like curve: slow random walk + dips ("pauses") every few thousand frames."""
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.standard_normal(n)) * 0.05
    x -= np.linspace(x[0], x[-1], n)
    for c in rng.integers(0, n, size=max(1, n // 3000)):
        w = int(rng.integers(30, 200))
        lo, hi = max(0, c - w), min(n, c + w)
        x[lo:hi] += rng.uniform(2.0, 8.0) * np.hanning(hi - lo)
    return x - 20.0 + shift + rng.standard_normal(n) * 0.01


def pause_audio(seed: int, seconds: float) -> np.ndarray:
    """int16 audio: noise bursts ("speech") separated by quiet gaps of random length."""
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    x = rng.standard_normal(n) * 3000.0
    env = np.ones(n)
    t = 0
    while t < n:
        t += int(rng.uniform(3.0, 25.0) * 16000)
        g = int(rng.uniform(0.2, 1.5) * 16000)
        env[t:t + g] = rng.uniform(0.002, 0.05)
        t += g
    return np.clip(x * env, -32768, 32767).astype(np.int16)


SEARCH_CASES = [
    # name, seed, n_frames, kwargs of BeamSearch
    ("search_default_7k", 1, 7000, dict(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)),
    ("search_default_30k", 2, 30000, dict(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)),
    ("search_default_120k", 3, 120000, dict(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)),
    ("search_short_1500", 4, 1500, dict(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)),
    ("search_beam3_step7_shift19.5", 5, 40000, dict(beam_size=3, ideal_segment_len=3000, step=7, len_reward_weight=2.0, energy_weight=3.0,
                                          min_len=1000, max_lookahead=9000)),
    ("search_energy_heavy_shift19.8", 6, 25000, dict(beam_size=5, ideal_segment_len=4000, step=10, len_reward_weight=0.5, energy_weight=4.0)),
    ("search_default_360k", 7, 360000, dict(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)),
]
CORE_CASES = [("core_150s", 11, 150.0, {}), ("core_400s", 12, 400.0, {}),
              ("core_300s_avg30", 13, 300.0, dict(average_segment_length=30.0, max_segment_len_sec=40))]


def main():
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
    psf = types.ModuleType("python_speech_features")
    psf.logfbank = psf_logfbank
    sys.modules.setdefault("python_speech_features", psf)
    sys.path.insert(0, "/root/reference")
    from speechcatcher import simple_endpointing as ref

    out = {"search": [], "core": []}
    for name, seed, n, kw in SEARCH_CASES:
        shift = float(name.split("shift")[1]) if "shift" in name else 0.0
        e = energy_curve(seed, n, shift)
        cuts = ref.BeamSearch(**kw).search(e, n)
        mine = CutSearchOracle(**kw).search(e, n)
        assert [tuple(c) for c in cuts] == mine, name
        out["search"].append(dict(name=name, seed=seed, n=n, shift=shift, kwargs=kw, cuts=[list(map(int, c)) for c in cuts]))
        print(name, len(cuts), "segments")
    for name, seed, secs, kw in CORE_CASES:
        a = pause_audio(seed, secs)
        segs = ref.segment_speech(a, 16000, **kw)
        mine = segment_speech_oracle(a, 16000, **kw)
        assert [tuple(s) for s in segs] == mine, (name, segs, mine)
        out["core"].append(dict(name=name, seed=seed, seconds=secs, kwargs=kw, segments=[list(map(int, s)) for s in segs]))
        print(name, segs)
    p = REPO / "tests" / "golden" / "endpointing.json"
    p.write_text(json.dumps(out, indent=1))
    print("wrote", p)


if __name__ == "__main__":
    main()
