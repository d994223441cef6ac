"""GPU: offline segmentation kernels against the CPU oracle and the reference-generated goldens."""
import json

import numpy as np
import pytest

from helpers import GOLDEN
from oracle.endpointing import gaussian_filter1d_reflect, psf_logfbank, segment_speech_oracle
from oracle.gen_golden_endpointing import pause_audio

pytestmark = pytest.mark.gpu
G = json.loads((GOLDEN / "endpointing.json").read_text())


@pytest.mark.parametrize("n", [100, 400, 401, 16000 * 3 + 77, 16000 * 70])
def test_energy_curve_matches_oracle(n):
    from speechcatcher_b200.simple_endpointing import smoothed_energy
    a = pause_audio(3, n / 16000.0)[:n]
    if n >= 16000:
        a[5000:9000] = 0                      # digital silence: exercises the zeros -> eps rule
    smooth, raw = smoothed_energy(a, 16000, return_raw=True)
    want_raw = psf_logfbank(a).sum(axis=-1) / 10.0
    assert raw.shape == want_raw.shape
    # fp64 on both sides; the FFT and the filter-bank sums differ only in summation order
    np.testing.assert_allclose(raw, want_raw, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(smooth, -gaussian_filter1d_reflect(want_raw, 20.0), rtol=1e-9, atol=1e-9)
    # given the device's own raw curve the smoothing follows scipy's summation order
    from scipy.ndimage import gaussian_filter1d
    np.testing.assert_allclose(smooth, -gaussian_filter1d(raw, sigma=20), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("case", G["core"], ids=lambda c: c["name"])
def test_segment_speech_matches_reference_golden(case):
    from speechcatcher_b200.simple_endpointing import segment_speech
    a = pause_audio(case["seed"], case["seconds"])
    assert [list(s) for s in segment_speech(a, 16000, **case["kwargs"])] == case["segments"]


def test_segment_speech_matches_live_oracle_on_fresh_audio():
    from speechcatcher_b200.simple_endpointing import segment_speech
    a = pause_audio(77, 240.0)
    assert segment_speech(a, 16000) == segment_speech_oracle(a, 16000)
