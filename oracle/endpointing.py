"""Oracle: offline segmentation of long files.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  speechcatcher/simple_endpointing.py:22-69   (BeamSearch: cost function + cut-point search)
  speechcatcher/simple_endpointing.py:72-98   (segment_speech_core: energy, smoothing, 180 s cap)
  speechcatcher/simple_endpointing.py:100-137 (segment_speech: default parameters)

Third-party arithmetic on this row:
  * `python_speech_features.logfbank` -- UNPINNED in the reference (requirements.txt) and ABSENT from
    this image.  `psf_logfbank` below restates the published algorithm of python_speech_features 0.6
    (sigproc.preemphasis / framesig / powspec, base.get_filterbanks / fbank / logfbank) in numpy fp64.
    ==> "parity unpinned" for the feature step: there is no installed copy to check it against.
  * `scipy.ndimage.gaussian_filter1d` -- present in this image (scipy); `gaussian_filter1d_reflect`
    restates it and tests/test_endpointing_cpu.py checks the restatement against scipy itself.
The cut-point search is pinned: oracle/gen_golden_endpointing.py runs the reference's own BeamSearch class
(imported from /root/reference with the two missing imports stubbed) on seeded energy curves and commits the
cuts under tests/golden/endpointing.json.
"""
from __future__ import annotations

import decimal
import math
from typing import List, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------- python_speech_features 0.6
def _round_half_up(x: float) -> int:
    return int(decimal.Decimal(x).quantize(decimal.Decimal("1"), rounding=decimal.ROUND_HALF_UP))


def psf_n_frames(n_samples: int, frame_len: int = 400, frame_step: int = 160) -> int:
    """sigproc.framesig: one frame for short signals, else 1 + ceil((n - len) / step) (zero padded tail)."""
    if n_samples <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def psf_filterbank_bins(nfilt: int = 26, nfft: int = 512, samplerate: int = 16000) -> np.ndarray:
    """base.get_filterbanks: FFT-bin edges of the triangular mel filters (HTK mel, floor to bins)."""
    hz2mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel2hz = lambda mel: 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    melpoints = np.linspace(hz2mel(0.0), hz2mel(samplerate / 2.0), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


def psf_filterbanks(nfilt: int = 26, nfft: int = 512, samplerate: int = 16000) -> np.ndarray:
    b = psf_filterbank_bins(nfilt, nfft, samplerate)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def psf_logfbank(signal: np.ndarray, samplerate: int = 16000, winlen: float = 0.025, winstep: float = 0.01,
                 nfilt: int = 26, nfft: int = 512, preemph: float = 0.97) -> np.ndarray:
    """base.logfbank(signal, samplerate, winlen, winstep) with its defaults (rectangular window,
    pre-emphasis 0.97, 26 filters, 512-point FFT, power spectrum / nfft, zeros replaced by eps)."""
    signal = np.asarray(signal)
    sig = np.append(signal[0], signal[1:] - preemph * signal[:-1]).astype(np.float64)
    frame_len = _round_half_up(winlen * samplerate)
    frame_step = _round_half_up(winstep * samplerate)
    n = psf_n_frames(len(sig), frame_len, frame_step)
    padlen = (n - 1) * frame_step + frame_len
    pad = np.concatenate((sig, np.zeros(padlen - len(sig))))
    idx = np.arange(frame_len)[None, :] + (np.arange(n) * frame_step)[:, None]
    frames = pad[idx]
    pspec = (1.0 / nfft) * np.square(np.absolute(np.fft.rfft(frames, nfft)))
    feat = np.dot(pspec, psf_filterbanks(nfilt, nfft, samplerate).T)
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    return np.log(feat)


# ----------------------------------------------------------------------------- scipy.ndimage.gaussian_filter1d
def gaussian_weights(sigma: float = 20.0, truncate: float = 4.0) -> np.ndarray:
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return w / w.sum()


def gaussian_filter1d_reflect(x: np.ndarray, sigma: float = 20.0, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage.gaussian_filter1d(x, sigma) with its defaults: mode='reflect' (d c b a | a b c d | d c b a),
    kernel radius int(4 sigma + 0.5)."""
    w = gaussian_weights(sigma, truncate)
    r = len(w) // 2
    n = len(x)
    idx = np.arange(-r, n + r)
    idx = np.mod(idx, 2 * n)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    return np.convolve(np.asarray(x, np.float64)[idx], w[::-1], mode="valid")


# ----------------------------------------------------------------------------- simple_endpointing.py
class CutSearchOracle:
    """simple_endpointing.py:22-69.  Quirks kept on purpose: the length reward is weighted twice
    (len_reward_factor already contains len_reward_weight, :32 and :40); `score_at_k` is the score of the LAST
    beam entry of the current generation (:50); a candidate is kept only if it improves on its parent (:56) while
    `expand` looks at the last entry's score (:60); a sequence that cannot be extended simply drops out (:46-62)."""

    def __init__(self, beam_size=10, ideal_segment_len=4000, max_lookahead=18000, min_len=2000, step=10,
                 len_reward_weight=1.0, energy_weight=1.0):
        self.beam_size, self.ideal, self.max_lookahead, self.min_len, self.step = \
            beam_size, ideal_segment_len, max_lookahead, min_len, step
        self.lw, self.ew = len_reward_weight, energy_weight
        self.factor = len_reward_weight / float(ideal_segment_len)

    def cost(self, seg_len: int, energy: float) -> float:
        reward = self.factor * (self.ideal - abs(self.ideal - float(seg_len)))
        return (self.lw * reward) + (self.ew * energy)

    def search(self, smoothed: Sequence[float], n_frames: int) -> List[Tuple[int, int]]:
        seqs = [([0], 0.0)]
        while True:
            cands = []
            expand = False
            worst = seqs[-1][1]
            for cuts, score in seqs:
                last = cuts[-1]
                for j in range(self.min_len, min(self.max_lookahead, n_frames - last - 1), self.step):
                    new = score + self.cost(j, smoothed[last + j])
                    if new > score:
                        cands.append((cuts + [last + j + 1], new))
                    if new > worst:
                        expand = True
            if not cands or not expand:
                break
            seqs = sorted(cands, key=lambda c: c[1], reverse=True)[: self.beam_size]   # stable, like the reference
        best = seqs[0][0] if seqs[0][0] != [0] else [0, n_frames]
        return list(zip(best[:-1], best[1:]))


def smoothed_energy(data: np.ndarray, samplerate: int = 16000) -> np.ndarray:
    """simple_endpointing.py:73-75: negated, Gaussian-smoothed sum of the 26 log filter-bank energies / 10."""
    fb = psf_logfbank(data, samplerate=samplerate, winlen=0.025, winstep=0.01)
    return gaussian_filter1d_reflect(fb.sum(axis=-1) / 10.0, sigma=20) * -1.0


def cap_segments(segments, max_segment_len_sec=180):
    """simple_endpointing.py:84-93: no segment longer than max_segment_len_sec."""
    cap = max_segment_len_sec * 100
    out = []
    for start, end in segments:
        while end - start > cap:
            out.append((start, start + cap))
            start += cap
        out.append((start, end))
    return out


def segment_speech_oracle(data, samplerate, average_segment_length=60.0, max_segment_len_sec=180, beam_size=10,
                          step=10, len_reward=40, len_reward_weight=12.0, energy_weight=1.0):
    """simple_endpointing.py:100-137 (`len_reward` is accepted and unused there too)."""
    search = CutSearchOracle(beam_size=beam_size, ideal_segment_len=int(average_segment_length * 100), step=step,
                             len_reward_weight=len_reward_weight, energy_weight=energy_weight)
    e = smoothed_energy(data, samplerate)
    return cap_segments(search.search(e, len(e)), max_segment_len_sec)
