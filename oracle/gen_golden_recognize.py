"""Generate tests/golden/recognize.json by running the REFERENCE's own `recognize` (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage (needs /root/reference):

    python -m oracle.gen_golden_recognize

`speechcatcher/speechcatcher.py` imports espnet_model_zoo, pyaudio, ffmpeg and (through simple_endpointing)
python_speech_features at module level; all four are absent from this image and are stubbed in sys.modules.  The
recogniser is oracle.scripted_backend.ScriptedSpeech2Text (a deterministic function of the samples a segment was
fed), the segmentation is forced per case; everything else -- finalize-iteration maths, chunk slicing, per-segment
reset, paragraph merging, capitalisation, timestamp conversion -- is the reference's own code
(speechcatcher.py:414-644).  Also pins `linear_interpolate_pos` (:323-358).
"""
from __future__ import annotations

import contextlib
import io
import json
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle.scripted_backend import ScriptedSpeech2Text  # noqa: E402

# name, seed, n_samples, chunk_length, forced segments (frame pairs) or None
CASES = [
    ("short_30s", 1, 30 * 16000 + 123, 8192, None),
    ("exact_multiple", 2, 8192 * 40, 8192, None),
    ("three_segments", 3, 200 * 16000 + 5000, 8192, [(0, 5421), (5421, 12007), (12007, 19000)]),
    ("boundary_in_last_10s_dropped", 4, 125 * 16000, 8192, [(0, 6000), (6000, 11800)]),
    ("many_segments_chunk4000", 5, 400 * 16000 + 1, 4000,
     [(0, 3011), (3011, 6142), (6142, 9143), (9143, 11834), (11834, 14835), (14835, 17836), (17836, 20757),
      (20757, 23758), (23758, 26759), (26759, 29760), (29760, 33333), (33333, 38000)]),
    ("tiny_2s", 6, 2 * 16000, 8192, None),
    ("empty_file", 7, 0, 8192, None),
    ("one_sample", 8, 1, 8192, None),
]
INTERP = [[3, 3, 3, 7, 9, 9, 12], [0, 0, 5, 5, 6], [1, 2, 3], [4, 4, 4, 4], [], [0], [2, 2, 10, 10, 10, 11, 30, 30]]


def case_audio(seed: int, n: int) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal(n) * 3000.0).clip(-32768, 32767).astype(np.int16)


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return m
    stub("espnet_model_zoo")
    stub("espnet_model_zoo.downloader", ModelDownloader=object)
    stub("pyaudio", paInt16=8)
    stub("ffmpeg")
    stub("python_speech_features", logfbank=None)
    sys.path.insert(0, "/root/reference")
    from speechcatcher import speechcatcher as ref
    return ref


def main():
    ref = import_reference()
    out = {"recognize": [], "interp": []}
    for name, seed, n, chunk, segs in CASES:
        a = case_audio(seed, n)
        ref.segment_speech = (lambda data, rate, _s=segs: list(_s or []))
        backend = ScriptedSpeech2Text()
        with contextlib.redirect_stdout(io.StringIO()):
            text, aux = ref.recognize(backend, a, 16000, chunk_length=chunk, num_processes=1, progress=False,
                                      quiet=True, decoder_impl="native")
        out["recognize"].append(dict(name=name, seed=seed, n=n, chunk=chunk, segments=segs, text=text, aux=aux,
                                     calls=backend.log))
        print(name, len(aux), "paragraphs", len(backend.log), "calls", repr(text[:60]))
    for lst in INTERP:
        out["interp"].append(dict(inp=lst, out=ref.linear_interpolate_pos(list(lst))))
    p = REPO / "tests" / "golden" / "recognize.json"
    p.write_text(json.dumps(out))
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
