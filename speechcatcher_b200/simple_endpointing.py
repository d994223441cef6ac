"""Offline segmentation of long files on the B200 path.

Same surface as the reference module `speechcatcher/simple_endpointing.py`: `BeamSearch(...).search(curve, n)`,
`segment_speech_core(data, samplerate, beam_search, max_segment_len_sec)` and `segment_speech(data, samplerate, ...)`
return the same lists of `(start, end)` frame pairs (100 frames per second).

Where the work runs: the energy curve (python_speech_features-style 26-filter log-fbank sum and the scipy Gaussian
smoothing, simple_endpointing.py:73-75) is computed in fp64 by two CUDA kernels (`sc_segment_energy`); the cut-point
search (:43-69) is sequential, data dependent and tiny, and runs on the host inside the C ABI (`sc_segment_search`).
There is no CPU fallback for the energy curve: without a CUDA device `segment_speech` raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ScSegmentParams


class BeamSearch:
    """Cut-point search with the reference's constructor (simple_endpointing.py:23-34)."""

    def __init__(self, beam_size=10, ideal_segment_len=4000, max_lookahead=18000, min_len=2000, step=10,
                 len_reward_weight=1.0, energy_weight=1.0):
        self.beam_size, self.ideal_segment_len, self.max_lookahead = int(beam_size), int(ideal_segment_len), int(max_lookahead)
        self.min_len, self.step = int(min_len), int(step)
        self.len_reward_weight, self.energy_weight = float(len_reward_weight), float(energy_weight)

    def search(self, smoothed_energy: Sequence[float], fbank_feat_len: int) -> List[Tuple[int, int]]:
        lib = _lib.load()
        curve = np.ascontiguousarray(smoothed_energy, np.float64)
        n = int(fbank_feat_len)
        if n > len(curve):
            raise ValueError(f"fbank_feat_len={n} exceeds the curve length {len(curve)}")
        p = ScSegmentParams(beam_size=self.beam_size, ideal_segment_len=self.ideal_segment_len,
                            max_lookahead=self.max_lookahead, min_len=self.min_len, step=self.step, reserved=0,
                            len_reward_weight=self.len_reward_weight, energy_weight=self.energy_weight)
        max_cuts = n // max(self.min_len, 1) + 4
        cuts = np.zeros(max_cuts, np.int64)
        n_cuts = C.c_int32()
        _lib.check(lib.sc_segment_search(curve.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(p),
                                         cuts.ctypes.data_as(C.POINTER(C.c_int64)), max_cuts, C.byref(n_cuts)),
                   "segment_search")
        c = cuts[: n_cuts.value].tolist()
        return list(zip(c[:-1], c[1:]))


def num_frames(n_samples: int) -> int:
    n = C.c_int64()
    _lib.check(_lib.load().sc_segment_num_frames(int(n_samples), C.byref(n)), "segment_num_frames")
    return n.value


def smoothed_energy(data: np.ndarray, samplerate: int = 16000, sigma: float = 20.0, device: str = "cuda:0",
                    return_raw: bool = False):
    """simple_endpointing.py:73-75 on the device (fp64).  `data`: the int16 samples of the whole file."""
    import torch
    if samplerate != 16000:
        raise ValueError("the segmenter is built for 16 kHz audio (the reference asserts rate == 16000, speechcatcher.py:427)")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available: the B200 segmenter cannot run (there is no CPU fallback)")
    data = np.asarray(data)
    if data.dtype != np.int16:
        raise TypeError(f"segment_speech takes the raw int16 samples like the reference, got {data.dtype}")
    lib = _lib.load()
    dev = torch.device(device)
    n = num_frames(len(data))
    with torch.cuda.device(dev):
        pcm = torch.from_numpy(np.ascontiguousarray(data)).to(dev, non_blocking=False)
        energy = torch.empty(n, dtype=torch.float64, device=dev)
        smooth = torch.empty(n, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev)
        _lib.check(lib.sc_segment_energy(C.c_void_p(pcm.data_ptr()), len(data), C.c_void_p(energy.data_ptr()),
                                         C.c_void_p(smooth.data_ptr()), n, float(sigma), C.c_void_p(st.cuda_stream)),
                   "segment_energy")
        out = smooth.cpu().numpy()
        if return_raw:
            return out, energy.cpu().numpy()
    return out


def cap_segments(segments, max_segment_len_sec=180):
    """No segment longer than max_segment_len_sec (simple_endpointing.py:84-93)."""
    cap = int(max_segment_len_sec * 100)
    out = []
    for start, end in segments:
        while end - start > cap:
            out.append((start, start + cap))
            start += cap
        out.append((start, end))
    return out


def segment_speech_core(data, samplerate, beam_search: BeamSearch, max_segment_len_sec=180, debug=False,
                        visual_debug=False, device: str = "cuda:0"):
    curve = smoothed_energy(data, samplerate, device=device)
    return cap_segments(beam_search.search(curve, len(curve)), max_segment_len_sec)


def segment_speech(data, samplerate, average_segment_length=60.0, max_segment_len_sec=180, beam_size=10, step=10,
                   len_reward=40, len_reward_weight=12.0, energy_weight=1.0, debug=False, visual_debug=False,
                   device: str = "cuda:0"):
    """simple_endpointing.py:100-137 (`len_reward` is accepted and unused, as in the reference)."""
    search = BeamSearch(beam_size=beam_size, ideal_segment_len=int(average_segment_length * 100), step=step,
                        len_reward_weight=len_reward_weight, energy_weight=energy_weight)
    return segment_speech_core(data, samplerate, search, max_segment_len_sec, debug, visual_debug, device=device)
