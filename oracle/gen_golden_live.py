"""Generate tests/golden/live.json by running the REFERENCE's own live-session code (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage (needs /root/reference):  python -m oracle.gen_golden_live

Drives `speechcatcher_server.SpeechRecognitionSession.process_audio_chunk` (speechcatcher_server.py:203-296) with the
scripted recogniser oracle.scripted_backend.ScriptedLive over seeded int16 chunk sequences (voiced / silent chunks,
empty chunks, Vosk `eof` / `reset` / config messages).  websockets, ffmpeg, pyaudio, espnet_model_zoo and
python_speech_features are absent from this image and stubbed; the per-session ffmpeg pipe is patched out (sessions are
fed 16 kHz int16 arrays, which `decode_audio` passes through, :177-179).
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

from oracle.gen_golden_recognize import import_reference  # noqa: E402
from oracle.scripted_backend import ScriptedLive  # noqa: E402

# name, seed, n_chunks, vosk, finalize_update_iters, max_partial_iters
CASES = [("plain_default", 1, 80, False, 7, 1024), ("vosk_default", 2, 80, True, 7, 1024),
         ("vosk_short_rule", 3, 60, True, 3, 1024), ("plain_max_iters", 4, 70, False, 5, 12),
         ("vosk_messages", 5, 50, True, 4, 1024)]


def chunk_script(seed: int, n: int, with_messages: bool):
    """List of chunks: int16 arrays (voiced noise or near-silence), empty arrays and, optionally, Vosk messages."""
    rng = np.random.default_rng(seed)
    out, voiced = [], True
    for i in range(n):
        if rng.random() < 0.15:
            voiced = not voiced
        r = rng.random()
        if with_messages and r < 0.06:
            out.append(rng.choice(['{"eof" : 1}', '{"reset" : 1}', '{"config" : {"sample_rate" : 16000}}']).item())
        elif r < 0.1:
            out.append(np.zeros(0, np.int16))
        else:
            amp = 3000.0 if voiced else 20.0
            out.append((rng.standard_normal(4096) * amp).clip(-32768, 32767).astype(np.int16))
    return out


def main():
    import_reference()
    for name in ("websockets",):
        sys.modules.setdefault(name, types.ModuleType(name))
    from speechcatcher import speechcatcher_server as srv
    srv.SpeechRecognitionSession.start_ffmpeg_process = lambda self, *a, **k: None
    out = []
    for name, seed, n, vosk, upd, mx in CASES:
        backend = ScriptedLive()
        with contextlib.redirect_stdout(io.StringIO()):
            sess = srv.SpeechRecognitionSession(backend, audio_format="s16le", finalize_update_iters=upd,
                                                max_partial_iters=mx, vosk_output_format=vosk)
            outputs = [sess.process_audio_chunk(c) for c in chunk_script(seed, n, "messages" in name)]
        n_final = sum(1 for _, _, f in backend.log if f)
        out.append(dict(name=name, seed=seed, n=n, vosk=vosk, finalize_update_iters=upd, max_partial_iters=mx,
                        outputs=outputs, calls=backend.log))
        print(name, len(outputs), "outputs", len(backend.log), "recogniser calls", n_final, "finalised")
    p = REPO / "tests" / "golden" / "live.json"
    p.write_text(json.dumps(out))
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
