"""Scripted stand-ins for the recogniser, used to pin the FILE-LEVEL glue without a model.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The text a segment yields is a deterministic function of the samples
the segment was fed, so any mistake in the call schedule (which samples, which call is final, when the stream is
reset, which call finalises everything) changes the output.  Two faces over the same core:
  * ScriptedSpeech2Text -- the protocol the reference's `recognize` drives (speechcatcher.py:592-644):
    `__call__(speech=, is_final=, finalize_all=, always_assemble_hyps=)` returning ESPnet-shaped 5-tuples, `reset()`;
  * ScriptedGroup       -- the StreamGroup protocol `speechcatcher_b200.recognize` drives (`push`, `last_plan`,
    `results`, `reset`), N independent streams.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List

import numpy as np

PIECES = ["▁hallo", "▁welt", "▁das", "▁ist", "▁ein", "▁test", "en", "ung", ".", "?", "!",
          "▁und", "▁so", "▁weiter", "▁Berlin", "▁ja"]


def _render(samples: np.ndarray, n_calls: int, finalize_all: bool):
    """Result list of a final call after `samples` were fed in `n_calls` calls."""
    n = len(samples)
    h = (int(np.abs(samples.astype(np.float64)).sum() * 1e4) + 31 * n + 7 * n_calls) % (2 ** 31)
    rng = np.random.default_rng(h)
    if not finalize_all and h % 7 == 0:
        return []                                  # an intermediate final call without a completed hypothesis
    n_tok = n // 16000
    ids = rng.integers(0, len(PIECES), size=n_tok)
    toks = [PIECES[i] for i in ids]
    if n_tok and h % 3 != 0:
        toks[-1] = "."                             # most segments end a sentence
    pos = np.sort(rng.integers(0, max(1, int(n / 16000 * 25)), size=n_tok)).tolist()
    text = "".join(toks).replace("▁", " ").strip()
    hyp = {"score": float(rng.standard_normal()), "n_calls": n_calls, "finalize_all": bool(finalize_all)}
    second = ("zweite hypothese", ["▁zweite"], [1], [0], {})
    return [(text, toks, [int(i) + 2 for i in ids], pos, hyp), second]


class _Stream:
    def __init__(self):
        self.reset()

    def reset(self):
        self.chunks: List[np.ndarray] = []
        self.n_calls = 0
        self.n_resets = getattr(self, "n_resets", 0) + 1

    def feed(self, chunk):
        self.chunks.append(np.asarray(chunk, np.float32))
        self.n_calls += 1

    def final_results(self, finalize_all):
        s = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.float32)
        return _render(s, self.n_calls, finalize_all)


class ScriptedSpeech2Text:
    def __init__(self):
        self.s = _Stream()
        self.log = []

    def reset(self):
        self.s.reset()

    def __call__(self, speech, is_final=False, finalize_all=False, always_assemble_hyps=True):
        self.log.append((len(speech), bool(is_final), bool(finalize_all)))
        self.s.feed(speech)
        return self.s.final_results(finalize_all) if is_final else []


class ScriptedGroup:
    def __init__(self, n_streams: int, max_seconds: float = 1e9, max_chunk: int = 1 << 30):
        self.n_streams, self.max_seconds, self.max_chunk = n_streams, max_seconds, max_chunk
        self.streams = [_Stream() for _ in range(n_streams)]
        self.token_list = None
        self.pushes = []

    def reset(self, streams=None):
        for s in (range(self.n_streams) if streams is None else streams):
            self.streams[s].reset()

    def push(self, ids, chunks, is_final):
        assert len(set(ids)) == len(ids)
        self.pushes.append([(int(s), len(c), bool(f)) for s, c, f in zip(ids, chunks, is_final)])
        for s, c in zip(ids, chunks):
            self.streams[s].feed(c)

    def last_plan(self, s):
        return SimpleNamespace(called=1)

    def results(self, s, is_final, finalize_all, token_list=None):
        assert is_final
        return self.streams[s].final_results(finalize_all)


class ScriptedLive:
    """Scripted recogniser for the LIVE session glue: partial texts whose length grows while the
    chunks carry signal and stalls on silent chunks, so the "length unchanged for N iterations" rule has something to
    look at.  Same protocol as ScriptedSpeech2Text; results are ESPnet-shaped 5-tuples."""

    def __init__(self, empty_until: int = 2):
        self.empty_until = empty_until
        self.reset()
        self.log = []

    def reset(self):
        self.voiced = 0
        self.calls = 0

    def __call__(self, speech, is_final=False, finalize_all=False, always_assemble_hyps=True):
        x = np.asarray(speech, np.float32)
        self.log.append((int(x.size), round(float(np.abs(x).sum()), 3), bool(is_final)))
        self.calls += 1
        if float(np.abs(x).max(initial=0.0)) > 0.01:
            self.voiced += 1
        if self.calls <= self.empty_until:
            return []
        toks = [PIECES[(3 * k + self.voiced) % len(PIECES)] for k in range(self.voiced)]
        text = "".join(toks).replace("▁", " ").strip()
        res = [(text, toks, list(range(2, 2 + len(toks))), [3 * k for k in range(len(toks))], {})]
        if is_final:
            self.voiced = 0        # the next utterance starts short again (the recogniser itself is not reset)
        return res
