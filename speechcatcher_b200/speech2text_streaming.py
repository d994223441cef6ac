"""Drop-in `Speech2TextStreaming` on the B200 CUDA path.

Same constructor, `__call__`, `reset`, `recognize`, `recognize_stream`, `n_best_hypotheses`,
`get_best_hypothesis` and `create_streaming_interface` as the reference class
(speechcatcher/speech2text_streaming.py:29-621).  Each instance is a one-stream view onto a
`StreamGroup`; pass `group=`/`stream_id=` to share one engine between many facades (what the
reference's server pool and segment loop do with N model copies, speechcatcher_server.py:331-357).

Compat superset: `always_assemble_hyps` is accepted (the reference CLI passes
it, speechcatcher.py:612) and results are ESPnet-shaped 5-tuples `(text, tokens, token_ids, token_pos,
hyp)` whose first three fields equal the reference's 3-tuple.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .model_files import load_tokenizer
from .stream_group import StreamGroup


class _SearchView:
    """What callers read off `speech2text.beam_search` in the reference (beam_search.py:266-327, 845-924)."""

    def __init__(self, facade):
        self._f = facade
        self.beam_size, self.use_bbd = facade.beam_size, facade.use_bbd
        self.weights = {"decoder": 1.0 - facade.ctc_weight, "ctc": facade.ctc_weight}
        self.scorers = {"decoder": "libscb200 decoder step (KV-cached)", "ctc": "libscb200 CTC prefix scorer"}
        self.block_size, self.hop_size, self.look_ahead, self.max_length = 40, 16, 16, 500
        self.sos = self.eos = facade.group.cfg.vocab - 1

    @property
    def process_idx(self):
        return self._f.beam_state[3] if self._f.beam_state is not None else 0

    def reset(self):
        self._f.reset()


class Speech2TextStreaming:
    def __init__(self, model_dir: Union[str, Path] = None, beam_size: int = 5, ctc_weight: float = 0.3,
                 device: str = "cuda", dtype: str = "float32", use_bbd: bool = False,
                 group: Optional[StreamGroup] = None, stream_id: int = 0, max_chunk: int = 8192,
                 max_seconds: float = 61.0):
        self._owns_group = group is None
        if group is None:
            if device == "cuda":
                device = "cuda:0"
            self._group_args = dict(model_dir=model_dir, n_streams=1, beam_size=beam_size, ctc_weight=ctc_weight,
                                    device=device, dtype=dtype, use_bbd=use_bbd, max_seconds=max_seconds)
            group = StreamGroup(max_chunk=max_chunk, **self._group_args)
            stream_id = 0
        self.group, self.stream_id = group, stream_id
        self._calls_since_reset = 0
        self.model_dir = Path(group.model_dir)
        self.beam_size, self.ctc_weight, self.use_bbd = group.beam_size, group.ctc_weight, group.use_bbd
        self.device, self.dtype = str(group.device), dtype
        self.mean, self.std = group.mean, group.std
        self.win_length, self.hop_length = 400, 160
        self.tokenizer, self.token_list = load_tokenizer(self.model_dir)      # speech2text_streaming.py:97-124
        # attributes the reference's callers / tests read (speech2text_streaming.py:62, 143-150): there is no
        # torch.nn.Module and no Python search object here, so `model` is the engine that holds the weights and
        # `beam_search` a read-only view of the search configuration
        self.model = group
        self.beam_search = _SearchView(self)
        self.beam_state = None
        self.processed_frames = 0
        self.frontend_states = None
        self.reset()

    def reset(self):
        self.beam_state = None
        self.processed_frames = 0
        self.frontend_states = None
        self._calls_since_reset = 0
        self.group.reset([self.stream_id])

    def __call__(self, speech: Union[np.ndarray, torch.Tensor], is_final: bool = False, finalize_all: bool = False,
                 always_assemble_hyps: bool = True) -> List[Tuple]:
        if isinstance(speech, torch.Tensor):
            speech = speech.detach().cpu().numpy()
        speech = np.asarray(speech, np.float32)
        if speech.ndim in (2, 3):
            return self._call_features(speech, is_final, finalize_all)
        if speech.ndim != 1:
            raise ValueError(f"speech must be a 1-D waveform, 2-D features or 3-D batched features, got {speech.ndim}-D")
        if len(speech) > self.group.max_chunk:
            # the engine's buffers are sized for a maximum chunk; a facade that owns its engine re-creates it with a
            # larger capacity, which is only possible while the stream holds no state (right after reset)
            if not self._owns_group or self._calls_since_reset > 0:
                raise ValueError(f"chunk of {len(speech)} samples exceeds the engine capacity max_chunk="
                                 f"{self.group.max_chunk}; construct with a larger max_chunk")
            need = (len(speech) + 159999) // 160000 * 160000
            self.ensure_capacity(max_chunk=need, max_seconds=need / 16000.0 + 2.0)
        self._calls_since_reset += 1
        self.group.push([self.stream_id], [speech], [is_final])
        plan = self.group.last_plan(self.stream_id)
        if not plan.called:
            return []                           # speech2text_streaming.py:431-433
        self.beam_state = self.group.beam(self.stream_id)
        return self.group.results(self.stream_id, is_final, finalize_all, self.token_list)

    def _call_features(self, feats: np.ndarray, is_final: bool, finalize_all: bool):
        """Pre-computed features (speech2text_streaming.py:438-450): 2-D input is normalised with the global statistics
        in numpy like the reference, 3-D (1, T, 80) input is used as it is; the frontend is skipped."""
        if feats.ndim == 3:
            if feats.shape[0] != 1:
                raise ValueError("batched feature input must have batch size 1 (one stream per facade)")
            feats = feats[0]
        elif self.mean is not None and self.std is not None:
            feats = ((feats - self.mean) / self.std).astype(np.float32)
        if len(feats) > self.group.max_feature_frames():
            if not self._owns_group or self._calls_since_reset > 0:
                raise ValueError(f"{len(feats)} feature frames exceed the engine capacity "
                                 f"{self.group.max_feature_frames()} per call; construct with a larger max_chunk")
            self.ensure_capacity(max_chunk=len(feats) * 160)
        self._calls_since_reset += 1
        self.group.push_features([self.stream_id], [feats], [is_final])
        self.beam_state = self.group.beam(self.stream_id)
        return self.group.results(self.stream_id, is_final, finalize_all, self.token_list)

    def ensure_capacity(self, n_streams: int = 1, max_seconds: float = 0.0, max_chunk: int = 0):
        """Re-create the owned engine if it is too small (streams, seconds per stream, samples per call).  Only a
        facade that owns its engine may do this, and only while the stream holds no state (right after reset): the
        file-level `recognize` uses it to run the segments of one file as concurrent streams."""
        g = self.group
        if g.n_streams >= n_streams and g.max_seconds >= max_seconds and g.max_chunk >= max_chunk:
            return
        if not self._owns_group:
            raise ValueError("this facade is a view onto a shared StreamGroup and cannot resize it")
        if self._calls_since_reset > 0:
            raise ValueError("the engine can only be resized right after reset()")
        args = {**self._group_args, "n_streams": max(n_streams, g.n_streams),
                "max_seconds": max(max_seconds, g.max_seconds)}
        new_chunk = max(max_chunk, g.max_chunk)
        StreamGroup.check_capacity(args["max_seconds"], new_chunk)      # raise BEFORE the old engine is given up
        old_chunk = g.max_chunk
        g.close()                                   # free the old workspace first: two engines may not fit side by side
        try:
            self.group = self.model = StreamGroup(max_chunk=new_chunk, **args)
        except Exception:
            # e.g. out of device memory: the facade must stay usable, so bring the previous engine back
            self.group = self.model = StreamGroup(max_chunk=old_chunk, **self._group_args)
            raise
        self._group_args = args

    def recognize(self, speech):
        self.reset()
        return self(speech, is_final=True)

    def recognize_stream(self, chunks):
        self.reset()
        results = None
        for i, chunk in enumerate(chunks):
            results = self(chunk, is_final=(i == len(chunks) - 1))
        return results if results is not None else []

    @property
    def n_best_hypotheses(self) -> int:
        return self.beam_size

    def get_best_hypothesis(self):
        if self.beam_state is None or not self.beam_state[0]:
            return None
        res = self.group.results(self.stream_id, True, True, self.token_list)
        return res[0] if res else None


def create_streaming_interface(model_dir, beam_size: int = 5, ctc_weight: float = 0.3, device: str = "cuda"):
    return Speech2TextStreaming(model_dir=model_dir, beam_size=beam_size, ctc_weight=ctc_weight, device=device)
