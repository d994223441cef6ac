"""Oracle: STFT -> power -> slaney mel -> log -> global MVN, with chunk buffering.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  speechcatcher/model/frontend/stft_frontend.py:14-154   (STFTFrontend)
  speechcatcher/speech2text_streaming.py:265-400          (normalize_features, apply_frontend)
  speechcatcher/model/checkpoint_loader.py:210-237        (load_normalization_stats)
and the published algorithm of torchaudio.functional.melscale_fbanks
(torchaudio 2.11, norm="slaney", mel_scale="slaney"), which the reference calls
at stft_frontend.py:73-81 and which is a third-party dependency.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

N_FFT = 512
HOP = 160
WIN = 400
N_MELS = 80
SR = 16000


def hann_window(win_length: int = WIN) -> torch.Tensor:
    """Periodic Hann window (torch.hann_window default; stft_frontend.py:68)."""
    n = torch.arange(win_length, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win_length)).to(torch.float32)


def _hz_to_mel_slaney(f: float) -> float:
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        mels = min_log_mel + math.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz_slaney(mels: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def mel_filterbank(n_freqs: int = N_FFT // 2 + 1, f_min: float = 0.0, f_max: float = SR / 2.0,
                   n_mels: int = N_MELS, sample_rate: int = SR) -> torch.Tensor:
    """Slaney-scale, slaney-normalised triangular filterbank, shape (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_slaney(f_min)
    m_max = _hz_to_mel_slaney(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def load_stats(stats_path) -> Tuple[np.ndarray, np.ndarray]:
    """mean/std in float64 (checkpoint_loader.py:210-237)."""
    st = np.load(stats_path)
    if "mean" in st:
        return st["mean"], st["std"]
    count = st["count"]
    mean = st["sum"] / count
    std = np.sqrt(np.maximum(st["sum_square"] / count - mean ** 2, 1e-10))
    return mean, std


def log_mel(wave: torch.Tensor, window: torch.Tensor, mel_fb: torch.Tensor) -> torch.Tensor:
    """(n,) waveform -> (frames, 80) log-mel, center=True reflect padding
    (stft_frontend.py:110-143)."""
    pad = N_FFT // 2
    x = torch.nn.functional.pad(wave.view(1, 1, -1), (pad, pad), mode="reflect").view(-1)
    n_frames = 1 + (x.numel() - N_FFT) // HOP
    frames = x.unfold(0, N_FFT, HOP)[:n_frames]
    left = (N_FFT - WIN) // 2
    w = torch.zeros(N_FFT, dtype=wave.dtype)
    w[left:left + WIN] = window
    spec = torch.fft.rfft(frames * w, n=N_FFT, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = torch.matmul(power, mel_fb)
    return torch.clamp(mel, min=1e-10).log()


class FrontendOracle:
    """One stream's waveform buffering / framing / trimming state machine
    (speech2text_streaming.py:278-400)."""

    def __init__(self, mean: Optional[np.ndarray], std: Optional[np.ndarray]):
        self.window = hann_window()
        self.mel_fb = mel_filterbank()
        self.mean, self.std = mean, std
        self.state = None  # {"waveform_buffer": tensor}

    def reset(self):
        self.state = None

    def __call__(self, speech: torch.Tensor, is_final: bool) -> Optional[torch.Tensor]:
        """Returns normalised features (T, 80) float32, or None when the call emits nothing."""
        prev = self.state
        if prev is not None:
            speech = torch.cat([prev["waveform_buffer"], speech], dim=0)
        if not speech.size(0) > WIN:                        # :306-319
            if is_final:
                speech = torch.cat([speech, torch.zeros(WIN - speech.size(0), dtype=speech.dtype)])
            else:
                self.state = {"waveform_buffer": speech.clone()}
                return None
        if is_final:                                        # :322-338
            to_process, buf = speech, None
        else:
            n_frames = (speech.size(0) - (WIN - HOP)) // HOP
            n_res = (speech.size(0) - (WIN - HOP)) % HOP
            to_process = speech[: (WIN - HOP) + n_frames * HOP]
            buf = speech[speech.size(0) - (WIN - HOP) - n_res:].clone()
        feats = log_mel(to_process.to(torch.float32), self.window, self.mel_fb)
        if self.mean is not None:                           # :355-358 numpy float64 round trip
            f = (feats.numpy() - self.mean) / self.std
            feats = torch.from_numpy(f).to(torch.float32)
        trim = math.ceil(math.ceil(WIN / HOP) / 2)          # :362  == 2
        n = feats.size(0)
        if is_final:
            if prev is not None and n > trim:
                feats = feats[trim:]
        else:
            if prev is None:
                if n > trim:
                    feats = feats[: n - trim]
            else:
                if n > 2 * trim:
                    feats = feats[trim: n - trim]
                else:                                        # :384-389
                    self.state = {"waveform_buffer": buf} if buf is not None else None
                    return None
        self.state = None if is_final else {"waveform_buffer": buf}
        return feats
