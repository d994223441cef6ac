"""Live sessions on the B200 path: the on-the-fly endpoint rule, the per-client session and a shared-engine pool
(SURVEY.md section 8(f), row N3).

Mirrors speechcatcher/speechcatcher_server.py:
  `SpeechRecognitionSession.process_audio_chunk` (:203-296: endpoint rule :252-268, output formatting :280-296),
  `format_vosk_partial` / `format_vosk_result` (:298-330), `Speech2TextPool` (:333-357)
and the microphone loop's variant of the rule, speechcatcher/speechcatcher.py:714-727.

What differs: the reference pre-loads `pool_size` copies of the model and steps every client on its own; here ONE
StreamGroup holds `pool_size` streams, `StreamPool.acquire()` hands out one-stream facades onto it, and
`process_many` advances any number of sessions with a single batched push (one pass through the CUDA path for all
of them).  Audio transcoding (the reference pipes every session through an ffmpeg process) is out of scope: sessions
take 16 kHz mono int16 PCM, as numpy arrays or s16le bytes.

Quirks kept: non-final results carry no committed tokens in the native decoder, so partial texts are
empty and the "length unchanged for N iterations" rule fires every N result-bearing iterations; after a finalised
utterance the stream is NOT reset (only a client `eof` / `reset` message resets the session's rule state).
"""
from __future__ import annotations

import json
from queue import Queue
from threading import Lock
from typing import List, Optional, Sequence

import numpy as np


class LiveEndpointer:
    """The rule of speechcatcher_server.py:252-268: finalise when the partial-text length has not changed over the last
    `finalize_update_iters` result-bearing iterations, or after more than `max_iters` of them.  `window` is the number
    of lengths compared (the server uses finalize_update_iters; the microphone loop compares the last 10 and has no
    iteration cap, speechcatcher.py:714-722: `LiveEndpointer(7, max_iters=None, window=10)`)."""

    def __init__(self, finalize_update_iters: int = 7, max_iters: Optional[int] = 1024, window: Optional[int] = None):
        self.finalize_update_iters, self.max_iters = finalize_update_iters, max_iters
        self.window = finalize_update_iters if window is None else window
        self.n_best_lens: List[int] = []

    def reset(self):
        self.n_best_lens = []

    def decide(self) -> bool:
        n = len(self.n_best_lens)
        if n < self.finalize_update_iters:
            return False
        if self.max_iters is not None and n > self.max_iters:
            self.n_best_lens = []
            return True
        if all(x == self.n_best_lens[-1] for x in self.n_best_lens[-1 * self.window:]):
            self.n_best_lens = []
            return True
        return False

    def observe(self, partial_len: int):
        self.n_best_lens.append(partial_len)


def format_vosk_partial(partial_text):
    return {"partial": partial_text}


def format_vosk_result(results, output_token_timestamps=True, real_timestamps=False):
    """speechcatcher_server.py:305-330: one entry per output token.  The reference fills in dummy times (0.1 s per
    token); `real_timestamps` uses the token positions the B200 path returns (encoder frame / 24 = seconds)."""
    words, text = [], ""
    if output_token_timestamps:
        tokens = results[0][1]
        pos = results[0][3] if real_timestamps and len(results[0]) >= 5 else None
        for idx, token in enumerate(tokens):
            start = pos[idx] / 24.0 if pos is not None else idx * 0.1
            words.append({"conf": 1.0, "start": start, "end": start + 0.1, "word": token.replace("▁", " ")})
            text += token
    return {"result": words, "text": text.replace("▁", " ").strip()}


class LiveSession:
    """One client's live session (SpeechRecognitionSession without the ffmpeg pipe)."""

    def __init__(self, speech2text, audio_format="s16le", finalize_update_iters=7, max_partial_iters=1024,
                 vosk_output_format=False, real_timestamps=False, reset_on_finalize=False, partials="reference"):
        if audio_format not in ("s16le", "pcm", "int16"):
            raise NotImplementedError("live sessions take 16 kHz mono s16le PCM; transcoding other container formats "
                                      "(the reference's per-session ffmpeg pipe) is outside the B200 path")
        self.speech2text = speech2text
        self.vosk_output_format, self.real_timestamps = vosk_output_format, real_timestamps
        # The reference never resets the recogniser after a finalised utterance, so its encoder memory and hypotheses
        # grow for the whole connection; the engine's per-stream capacity (max_seconds) is finite and a push beyond it
        # fails loudly.  reset_on_finalize=True starts every utterance from a clean stream (a deliberate deviation
        # for long-running connections); False is the reference's behaviour.
        self.reset_on_finalize = reset_on_finalize
        # partials="reference": non-final calls return what the reference returns -- ended hypotheses with NO committed
        # tokens, so partial texts are empty and the rule degenerates to "every N result-bearing calls".
        # partials="best": the partial is the text of the best running hypothesis (a read of the current beam, no
        # state change), which gives the "length unchanged for N iterations" rule real lengths to look at.
        if partials not in ("reference", "best"):
            raise ValueError("partials must be 'reference' or 'best'")
        self.partials = partials
        self.vosk_sample_rate = self.decoder_sample_rate = 16000
        self.rule = LiveEndpointer(finalize_update_iters, max_partial_iters)

    def reset(self):
        self.rule.reset()

    # -- the three stages of process_audio_chunk, split so that many sessions can share one batched push ----------
    def _prepare(self, audio_chunk):
        """-> (early_return | None, float16 samples, finalize_iteration, client_forced_finalize)."""
        forced = False
        if isinstance(audio_chunk, str):                                    # :207-224 configuration messages
            if not self.vosk_output_format:
                return "", None, False, False
            if audio_chunk in ('{"eof" : 1}', '{"reset" : 1}'):
                forced = True
                audio_chunk = np.zeros(1000, dtype=np.int16)
            else:
                try:
                    cfg = json.loads(audio_chunk)
                    if "config" in cfg and "sample_rate" in cfg["config"]:
                        self.vosk_sample_rate = int(cfg["config"]["sample_rate"])
                        if self.vosk_sample_rate != self.decoder_sample_rate:
                            raise NotImplementedError("resampling is outside the B200 path: send 16 kHz PCM")
                except json.JSONDecodeError:
                    pass
                return format_vosk_partial(""), None, False, False
        data = audio_chunk if isinstance(audio_chunk, np.ndarray) and audio_chunk.dtype == np.int16 \
            else np.frombuffer(audio_chunk, dtype="int16")
        if data.size == 0:
            return (format_vosk_partial("") if self.vosk_output_format else ""), None, False, False
        data = data.astype(np.float16) / 32767.0                            # :245 (the server feeds fp16-rounded samples)
        fin = self.rule.decide()
        if forced:
            fin = True
        return None, data, fin, forced

    def _finish(self, results, fin, forced):
        if forced:
            self.reset()
        if fin and self.reset_on_finalize:
            self.speech2text.reset()
        if results is not None and len(results) > 0:
            nbests0 = results[0][0]
            if fin:
                if len(nbests0) >= 1:
                    if nbests0[-1] not in ".!?":
                        nbests0 += "."
                    nbests0 += "\n"
            else:
                self.rule.observe(len(nbests0))
            if self.vosk_output_format:
                return format_vosk_result(results, real_timestamps=self.real_timestamps) if fin \
                    else format_vosk_partial(nbests0)
            return nbests0
        return ""

    def process_audio_chunk(self, audio_chunk, is_final=False, save_debug_wav=False, debug=False):
        early, data, fin, forced = self._prepare(audio_chunk)
        if early is not None:
            return early
        results = self.speech2text(speech=data, is_final=fin)
        if self.partials == "best" and not fin:
            best = self.speech2text.get_best_hypothesis()
            results = [best] if best is not None else []
        return self._finish(results, fin, forced)


def process_many(sessions: Sequence[LiveSession], audio_chunks: Sequence) -> List:
    """One step of many live sessions with a single batched push.  Every session's `speech2text` must be a facade
    onto the same StreamGroup (e.g. from one StreamPool).  Returns each session's `process_audio_chunk` output."""
    if not sessions:
        return []
    group = sessions[0].speech2text.group
    out: List = [None] * len(sessions)
    todo = []
    for i, (s, chunk) in enumerate(zip(sessions, audio_chunks)):
        if s.speech2text.group is not group:
            raise ValueError("process_many needs sessions that share one StreamGroup")
        early, data, fin, forced = s._prepare(chunk)
        if early is not None:
            out[i] = early
        else:
            todo.append((i, s, data, fin, forced))
    if todo:
        ids = [s.speech2text.stream_id for _, s, _, _, _ in todo]
        if len(set(ids)) != len(ids):
            raise ValueError("process_many got two sessions on the same stream")
        group.push(ids, [np.asarray(d, np.float32) for _, _, d, _, _ in todo], [f for _, _, _, f, _ in todo])
        for i, s, _, fin, forced in todo:
            f = s.speech2text
            f._calls_since_reset += 1
            if group.last_plan(f.stream_id).called:
                f.beam_state = group.beam(f.stream_id)
                if s.partials == "best" and not fin:
                    results = group.results(f.stream_id, True, True, f.token_list)[:1]
                else:
                    results = group.results(f.stream_id, fin, False, f.token_list)
            else:
                results = []
            out[i] = s._finish(results, fin, forced)
    return out


class StreamPool:
    """Speech2TextPool (speechcatcher_server.py:333-357) over ONE engine: `pool_size` streams of a shared StreamGroup
    instead of `pool_size` model copies.  `acquire()` returns a Speech2TextStreaming facade (or None when every stream
    is taken), `release()` puts it back -- after resetting the stream, which the reference forgets (its next client
    inherits the previous client's hypotheses)."""

    def __init__(self, model_dir, device="cuda:0", beam_size=3, pool_size=8, dtype="float32", max_seconds=61.0,
                 max_chunk=8192, group=None, **group_kw):
        from .speech2text_streaming import Speech2TextStreaming
        from .stream_group import StreamGroup
        self.pool_size = pool_size
        self.group = group if group is not None else StreamGroup(
            model_dir, n_streams=pool_size, beam_size=beam_size, device=device, dtype=dtype,
            max_seconds=max_seconds, max_chunk=max_chunk, **group_kw)
        self.pool: Queue = Queue(maxsize=pool_size)
        self.lock = Lock()
        for k in range(pool_size):
            self.pool.put(Speech2TextStreaming(group=self.group, stream_id=k, dtype=dtype))

    def acquire(self):
        with self.lock:
            if self.pool.empty():
                return None
            return self.pool.get()

    def release(self, model):
        with self.lock:
            model.reset()
            self.pool.put(model)
