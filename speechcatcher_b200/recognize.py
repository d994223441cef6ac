"""File-level recognition on the B200 path: segments of one long file run as concurrent streams of one engine
(SURVEY.md section 8(f), row N1).

Mirrors the reference's caller of the hot path, speechcatcher/speechcatcher.py:
  `recognize` (:414-566), `recognize_segment` (:570-589), `batch_recognize_inner_loop` (:592-644), `is_completed`
  (:316), `upperCaseFirstLetter` (:309), `linear_interpolate_pos` (:323-358)
with the same arguments and the same `(complete_text, auxiliary_info)` result.  What differs is where the segments run:
the reference forks `num_processes` workers, each holding a copy of the model and decoding its segments one after the
other; here up to `n_streams` segments advance in lock step through ONE StreamGroup (one batched push per chunk index),
so `num_processes` only asks for at least that many streams.  A stream is reset after its segment's final call,
exactly like `speech2text_global.reset()` at :609-610, and a freed stream immediately picks up the next waiting segment.

Segment -> call schedule (speechcatcher.py:431-446, 570-589), chunk length c, n samples:
  max_i = n // c + 1; a segment boundary at frame f (10 ms) finalises at iteration ceil((f / 100 * rate - c) / c);
  segment (start, end] covers iterations start+1 .. end; call i takes samples [i c, min((i + 1) c, n)), is final when
  i == end and outputs every hypothesis (`finalize_all`) only when i == max_i.  The very last call is therefore an
  EMPTY final chunk.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

espnet_input_factor = 24.0          # encoder frames per second of token positions (speechcatcher.py:48)


def upperCaseFirstLetter(utterance_text: str) -> str:
    if len(utterance_text) > 0 and utterance_text[0].islower():
        utterance_text = utterance_text[0].upper() + utterance_text[1:]
    return utterance_text


def is_completed(utterance: str) -> bool:
    return utterance.endswith((".", "?", "!"))


def linear_interpolate_pos(input_list_in: Sequence[float]) -> List[float]:
    """speechcatcher.py:323-358: spread runs of equal token positions linearly between the previous distinct position
    (0 before the first run) and the run's value; the last element of a run keeps the value."""
    vals = [0] + list(input_list_in)
    out: List[float] = []
    i = 0
    while i < len(vals):
        cur = float(vals[i])
        g0 = i
        while i < len(vals) and vals[i] == cur:
            i += 1
        g1 = i - 1
        prev = 0.0 if g0 == 0 else vals[g0 - 1]
        # quirk kept: for the first run (g0 == 0, only longer than the boundary zero when the positions start with
        # zeros) the reference's difference term indexes input_list[-1], i.e. the LAST position of the list
        base = vals[g0 - 1]
        out.extend(prev + (g1 - j) / (g1 - g0 + 1) * (cur - base) for j in range(g0, g1))
        out.append(float(vals[g1]))
    return out[1:]


@dataclass
class SegmentPlan:
    """Finalize schedule of one file (speechcatcher.py:431-446)."""
    max_i: int
    finalize_iters: List[int]                       # [-1, ..., max_i]: segment k covers iterations (f[k], f[k+1]]
    seconds: List[Tuple[float, float]]              # (start, end) of every segment in seconds

    @property
    def n_segments(self) -> int:
        return len(self.finalize_iters) - 1

    def calls(self, k: int, n_samples: int, chunk_length: int):
        """[(sample_start, sample_end, is_final, finalize_all)] of segment k (speechcatcher.py:574-586)."""
        start, end = self.finalize_iters[k], self.finalize_iters[k + 1]
        out = []
        for i in range(start + 1, end + 1):
            s, e = i * chunk_length, min((i + 1) * chunk_length, n_samples)
            out.append((min(s, n_samples), max(e, min(s, n_samples)), i == end, i == self.max_i))
        return out


def plan_segments(n_samples: int, rate: int, segments: Sequence[Tuple[int, int]], chunk_length: int) -> SegmentPlan:
    """Segment ends (10 ms frames) -> finalize iterations and segment times.  A boundary closer than 10 s to the end
    of the file is dropped so the last segment is at least 10 s long (speechcatcher.py:433-446)."""
    n_frames = (n_samples / rate) * 100.0
    ends = [seg[1] for seg in segments if seg[1] < n_frames - 1000.0]
    max_i = (n_samples // chunk_length) + 1
    secs = [0] + [f / 100.0 for f in ends] + [n_samples / rate]
    iters = [-1] + [math.ceil((((f / 100.0) * rate) - chunk_length) / chunk_length) for f in ends] + [max_i]
    return SegmentPlan(max_i, iters, list(zip(secs[:-1], secs[1:])))


def _engine_of(speech2text, want_streams: int, need_seconds: float, chunk_length: int):
    """(group, usable stream ids, token_list).  A facade that owns its engine grows it to `want_streams` streams and
    to the longest segment; a facade onto a shared group uses only its own stream; a StreamGroup uses all of its."""
    group = getattr(speech2text, "group", None)
    if group is None:                                   # a StreamGroup (or anything with its push/results protocol)
        group, streams, token_list = speech2text, list(range(speech2text.n_streams)), getattr(speech2text, "token_list", None)
    else:
        if getattr(speech2text, "_owns_group", False):
            speech2text.ensure_capacity(n_streams=want_streams, max_seconds=need_seconds, max_chunk=chunk_length)
            group = speech2text.group
            streams = list(range(group.n_streams))
        else:
            streams = [speech2text.stream_id]
        token_list = speech2text.token_list
    have = getattr(group, "max_seconds", None)
    if have is not None and have < need_seconds:
        raise ValueError(f"the longest segment needs an engine capacity of {need_seconds:.0f} s per stream, "
                         f"this StreamGroup was built with max_seconds={have:.0f}")
    if getattr(group, "max_chunk", chunk_length) < chunk_length:
        raise ValueError(f"chunk_length={chunk_length} exceeds the engine's max_chunk={group.max_chunk}")
    return group, streams, token_list


def recognize(speech2text, raw_speech_data, rate, chunk_length=8192, num_processes=1, progress=True, quiet=False,
              status=None, decoder_impl="native", segments: Optional[Sequence[Tuple[int, int]]] = None,
              segmenter: Optional[Callable] = None, on_step: Optional[Callable[[int, int], None]] = None,
              shard: Optional[Tuple[int, int]] = None):
    """Transcribe the int16 samples of one file; returns `(complete_text, auxiliary_info)` like speechcatcher.py:414.

    speech2text: a `speechcatcher_b200.Speech2TextStreaming` (or a `StreamGroup`).  `segments` overrides the offline
    segmentation (list of (start, end) frame pairs as `segment_speech` returns them); `segmenter(data, rate)` replaces
    `speechcatcher_b200.simple_endpointing.segment_speech`.  `on_step(done_calls, total_calls)` reports progress;
    `status.publish_status(str)` is called like the reference's status thread (every 10 calls).

    `shard=(rank, world)`: one process per GPU; every rank plans the same segments, decodes segments
    k with k % world == rank on its own engine -- no data-path collective -- and the per-segment results are exchanged
    once at the end (`sharding.gather_results`, NCCL or gloo), so every rank returns the complete transcript.
    """
    if decoder_impl != "native":
        raise ValueError("the B200 path implements the native decoder only (decoder_impl='native')")
    assert rate == 16000
    raw = np.asarray(raw_speech_data)
    speech = raw.astype(np.float32) / 32768.0                        # native decoder scaling (:421)
    n = len(speech)
    if segments is None:
        segments = []
        if n > 60.0 * rate:                                            # files longer than a minute are segmented
            if segmenter is None:
                from .simple_endpointing import segment_speech as segmenter
            segments = segmenter(raw, rate)
    plan = plan_segments(n, rate, segments, chunk_length)
    K = plan.n_segments
    calls = [plan.calls(k, n, chunk_length) for k in range(K)]
    longest = max((c[-1][1] - c[0][0]) for c in calls if c) if any(calls) else 0
    group, streams, token_list = _engine_of(speech2text, max(1, min(int(num_processes), K)),
                                            longest / rate + 2.0, chunk_length)

    done_calls = 0
    # per-segment result: [text, tokens, positions, hyp] (batch_recognize_inner_loop's return value)
    seg_out: List[list] = [["", [], [], {}] for _ in range(K)]
    mine = list(range(K)) if shard is None else [k for k in range(K) if k % shard[1] == shard[0]]
    total_calls = sum(len(calls[k]) for k in mine)
    waiting = list(mine)
    active = {}                                                        # stream id -> [segment, next call index]
    for s in streams:
        group.reset([s])
    while waiting or active:
        for s in streams:                                              # free streams pick up waiting segments
            if s not in active and waiting:
                k = waiting.pop(0)
                if calls[k]:
                    active[s] = [k, 0]
        if not active:
            continue
        ids = sorted(active)
        batch = [calls[active[s][0]][active[s][1]] for s in ids]
        group.push(ids, [speech[a:b] for a, b, _, _ in batch], [f for _, _, f, _ in batch])
        for s, (a, b, fin, fin_all) in zip(ids, batch):
            k, ci = active[s]
            if fin:
                # quiet/progress branch of :612-617: only the final call's first result is kept
                results = group.results(s, True, fin_all, token_list) if group.last_plan(s).called else []
                if results:
                    r = results[0]
                    seg_out[k] = [r[0], r[1], r[-2], r[-3]]
                group.reset([s])                                       # :609-610
                del active[s]
            else:
                active[s][1] = ci + 1
            done_calls += 1
            if on_step is not None:
                on_step(done_calls, total_calls)
            if status is not None and done_calls % 10 == 0:
                status.publish_status(f"Decoding progress: {done_calls / max(total_calls, 1) * 100.0:.2f}%")
    if shard is not None and shard[1] > 1:
        from .sharding import gather_results
        for k, v in gather_results({k: seg_out[k] for k in mine}).items():
            seg_out[k] = v
    return merge_paragraphs(seg_out, plan.seconds)


def merge_paragraphs(paragraphs_raw: Sequence[Sequence], seconds: Sequence[Tuple[float, float]]):
    """speechcatcher.py:499-566: a segment starts a new paragraph only if the previous segment's text ends a sentence
    (endpointer and model agree); otherwise it is appended to the open paragraph.  Token positions become seconds:
    segment start + position / 24."""
    texts = [p[0] for p in paragraphs_raw]
    toks = [p[1] for p in paragraphs_raw]
    poss = [p[2] for p in paragraphs_raw]
    n_tok = len(list(itertools.chain(*toks)))
    n_pos = sum(len(p) for p, _ in zip(poss, seconds))
    assert n_tok == n_pos                                               # :513
    merged = [texts[0]] if texts else []
    aux = []
    if texts:
        s0 = seconds[0]
        aux.append({"start": s0[0], "end": s0[1], "text": texts[0], "tokens": toks[0],
                    "token_timestamps": [s0[0] + float(t) / espnet_input_factor for t in poss[0]]})
    for prev, text, tk, ps, se in zip(texts[:-1], texts[1:], toks[1:], poss[1:], seconds[1:]):
        stamps = [se[0] + float(t) / espnet_input_factor for t in ps]
        assert len(tk) == len(stamps)
        if is_completed(prev):
            text = upperCaseFirstLetter(text)
            merged.append(text)
            aux.append({"start": se[0], "end": se[1], "text": text, "tokens": tk, "token_timestamps": stamps})
        else:
            merged[-1] += " " + text
            aux[-1]["end"] = se[1]
            aux[-1]["text"] += " " + text
            aux[-1]["tokens"].extend(tk)
            aux[-1]["token_timestamps"].extend(stamps)
    return "\n\n".join(merged) + "\n", aux
