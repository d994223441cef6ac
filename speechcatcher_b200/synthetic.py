"""Synthetic model directories and audio for the named architectures.

There is no network, so checkpoints cannot be downloaded: every parity run and
benchmark uses random-init weights of the named architecture, written in the
ESPnet checkpoint layout that the reference loads
(reference: speechcatcher/speech2text_streaming.py:157-250 `_load_model`,
tests/test_speech2text_streaming.py:19-62 dummy-model-dir fixture).

Weights are drawn with numpy's PCG64 (platform independent) using PyTorch's
default init distributions, so the same `(arch, seed)` gives bit-identical
tensors on the build container and on the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from pathlib import Path
from typing import Dict

import numpy as np


@dataclass(frozen=True)
class Arch:
    """Architecture hyper-parameters read from config.yaml by the reference.

    `linear_units` (2048), block 40 / hop 16 / look-ahead 16 and vocab 1024 are
    NOT configurable in the reference (speech2text_streaming.py:210-232,
    beam_search.py:281-292), so they are constants here too.
    """
    name: str
    d_model: int = 256
    enc_heads: int = 4
    enc_layers: int = 12
    dec_heads: int = 4
    dec_layers: int = 6
    vocab: int = 1024
    ffn: int = 2048
    n_mels: int = 80


# XL is documented; M = loader defaults; L = declared assumption.
ARCHS: Dict[str, Arch] = {
    "xl": Arch("xl", 256, 8, 30, 8, 14),
    "l": Arch("l", 256, 8, 18, 8, 8),
    "m": Arch("m", 256, 4, 12, 4, 6),
    # reduced-depth variants used by fast parity tests (same kernels, fewer layers)
    "xl_d4": Arch("xl_d4", 256, 8, 4, 8, 3),
    "m_d2": Arch("m_d2", 256, 4, 2, 4, 2),
}


def _linear(rng, out_f, in_f, fan_in=None):
    fan_in = fan_in or in_f
    bound = 1.0 / math.sqrt(fan_in)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def make_state_dict(arch: Arch, seed: int = 0, sharpen: float = 1.0, eos_bias: float = 0.0) -> Dict[str, np.ndarray]:
    """Random-init weights keyed like the reference's state_dict.

    `sharpen` multiplies `decoder.output_layer.weight` and `ctc.ctc_lo.weight`
    (SURVEY.md 8(d) "sharpened" weight set: wider decision margins).  `eos_bias` is added to the decoder's
    output bias of <sos/eos> (the last token): with plain random weights a hypothesis practically never ends, so
    segment ends inside a file (is_final without finalize_all) would return nothing; ~6-8 makes them end regularly.
    """
    rng = np.random.default_rng(seed)
    D, F, V = arch.d_model, arch.ffn, arch.vocab
    sd: Dict[str, np.ndarray] = {}

    def put_linear(prefix, out_f, in_f):
        w, b = _linear(rng, out_f, in_f)
        sd[prefix + ".weight"] = w
        sd[prefix + ".bias"] = b

    def put_norm(prefix):
        # a non-trivial affine so that LayerNorm scale/shift are exercised
        sd[prefix + ".weight"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        sd[prefix + ".bias"] = (0.1 * rng.standard_normal(D)).astype(np.float32)

    def put_mha(prefix):
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            put_linear(f"{prefix}.{n}", D, D)

    # conv2d subsampling (reference: model/encoder/subsampling.py:52-69)
    bound = 1.0 / math.sqrt(9)
    sd["encoder.embed.conv.0.weight"] = rng.uniform(-bound, bound, (D, 1, 3, 3)).astype(np.float32)
    sd["encoder.embed.conv.0.bias"] = rng.uniform(-bound, bound, (D,)).astype(np.float32)
    bound = 1.0 / math.sqrt(D * 9)
    sd["encoder.embed.conv.2.weight"] = rng.uniform(-bound, bound, (D, D, 3, 3)).astype(np.float32)
    sd["encoder.embed.conv.2.bias"] = rng.uniform(-bound, bound, (D,)).astype(np.float32)
    f2 = ((arch.n_mels - 3) // 2 + 1 - 3) // 2 + 1  # 19 for 80 mel bins
    put_linear("encoder.embed.out", D, D * f2)
    for l in range(arch.enc_layers):
        p = f"encoder.encoders.{l}"
        put_mha(p + ".self_attn")
        put_linear(p + ".feed_forward.w_1", F, D)
        put_linear(p + ".feed_forward.w_2", D, F)
        put_norm(p + ".norm1")
        put_norm(p + ".norm2")
    put_norm("encoder.after_norm")

    sd["decoder.embed.0.weight"] = rng.standard_normal((V, D)).astype(np.float32)
    put_norm("decoder.after_norm")
    for l in range(arch.dec_layers):
        p = f"decoder.decoders.{l}"
        put_mha(p + ".self_attn")
        put_mha(p + ".src_attn")
        put_linear(p + ".feed_forward.w_1", F, D)
        put_linear(p + ".feed_forward.w_2", D, F)
        put_norm(p + ".norm1")
        put_norm(p + ".norm2")
        put_norm(p + ".norm3")
    put_linear("decoder.output_layer", V, D)
    put_linear("ctc.ctc_lo", V, D)
    if sharpen != 1.0:
        sd["decoder.output_layer.weight"] = sd["decoder.output_layer.weight"] * np.float32(sharpen)
        sd["ctc.ctc_lo.weight"] = sd["ctc.ctc_lo.weight"] * np.float32(sharpen)
    if eos_bias != 0.0:
        sd["decoder.output_layer.bias"][V - 1] += np.float32(eos_bias)
    return sd


def make_feats_stats(n_mels: int = 80, seed: int = 0):
    """Global mean/variance stats in the `count/sum/sum_square` layout
    (reference: model/checkpoint_loader.py:210-237)."""
    rng = np.random.default_rng(seed)
    mean = rng.normal(-8.0, 1.0, n_mels)
    var = rng.uniform(2.0, 6.0, n_mels)
    count = np.float64(1.0e6)
    return {
        "count": np.array(count),
        "sum": (mean * count).astype(np.float64),
        "sum_square": ((var + mean ** 2) * count).astype(np.float64),
    }


def make_model_dir(path, arch="xl", seed: int = 0, sharpen: float = 1.0, eos_bias: float = 0.0) -> Path:
    """Write `valid.acc.best.pth`, `config.yaml`, `feats_stats.npz` under `path`."""
    import torch
    import yaml

    a = ARCHS[arch] if isinstance(arch, str) else arch
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    sd = {k: torch.from_numpy(v.copy()) for k, v in make_state_dict(a, seed, sharpen, eos_bias).items()}
    torch.save({"model": sd}, path / "valid.acc.best.pth")
    cfg = {
        "encoder_conf": {"output_size": a.d_model, "attention_heads": a.enc_heads,
                         "num_blocks": a.enc_layers},
        "decoder_conf": {"attention_heads": a.dec_heads, "num_blocks": a.dec_layers},
        "frontend_conf": {"n_fft": 512, "hop_length": 160, "win_length": 400},
        "arch": asdict(a),
    }
    with open(path / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    np.savez(path / "feats_stats.npz", **make_feats_stats(a.n_mels, seed))
    return path


def synth_audio(stream: int, n_samples: int, kind: str = "noise") -> np.ndarray:
    """Synthetic 16 kHz audio for stream `stream` (SURVEY.md 8(d) "Audio").

    noise: white noise sigma 0.05.  tones: noise + 3 random sinusoids.
    """
    rng = np.random.default_rng(1000 + stream)
    x = rng.standard_normal(n_samples).astype(np.float32) * np.float32(0.05)
    if kind == "tones":
        t = np.arange(n_samples, dtype=np.float64) / 16000.0
        for _ in range(3):
            f = rng.uniform(100.0, 4000.0)
            x = x + (0.1 * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))).astype(np.float32)
    return x
