"""CPU: offline segmentation -- the oracle against the reference-generated goldens and scipy,
and the product's host-side cut search / frame maths (C ABI, no GPU needed) against both."""
import ctypes as C
import json

import numpy as np
import pytest

from helpers import GOLDEN
from oracle.endpointing import (CutSearchOracle, gaussian_filter1d_reflect, psf_filterbank_bins, psf_logfbank,
                                psf_n_frames, segment_speech_oracle)
from oracle.gen_golden_endpointing import energy_curve, pause_audio

G = json.loads((GOLDEN / "endpointing.json").read_text())


@pytest.mark.parametrize("case", G["search"], ids=lambda c: c["name"])
def test_oracle_search_matches_reference_golden(case):
    e = energy_curve(case["seed"], case["n"], case["shift"])
    got = CutSearchOracle(**case["kwargs"]).search(e, case["n"])
    assert [list(c) for c in got] == case["cuts"]


@pytest.mark.parametrize("case", G["search"], ids=lambda c: c["name"])
def test_product_host_search_matches_reference_golden(case):
    from speechcatcher_b200.simple_endpointing import BeamSearch
    e = energy_curve(case["seed"], case["n"], case["shift"])
    got = BeamSearch(**case["kwargs"]).search(e, case["n"])
    assert [list(c) for c in got] == case["cuts"]


def test_product_host_search_random_parameters_match_oracle():
    from speechcatcher_b200.simple_endpointing import BeamSearch
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(200, 20000))
        kw = dict(beam_size=int(rng.integers(1, 8)), ideal_segment_len=int(rng.integers(100, 3000)),
                  max_lookahead=int(rng.integers(500, 6000)), min_len=int(rng.integers(0, 600)),
                  step=int(rng.integers(1, 30)), len_reward_weight=float(rng.uniform(0.1, 5)),
                  energy_weight=float(rng.uniform(0.1, 5)))
        e = energy_curve(100 + trial, n, shift=20.0 - float(rng.uniform(0.0, 2.0)))
        assert BeamSearch(**kw).search(e, n) == CutSearchOracle(**kw).search(e, n), (trial, kw)


def test_oracle_gaussian_matches_scipy():
    from scipy.ndimage import gaussian_filter1d
    rng = np.random.default_rng(0)
    for n, sigma in ((5000, 20), (100, 20), (37, 20), (1000, 3.5)):
        x = rng.standard_normal(n)
        np.testing.assert_allclose(gaussian_filter1d_reflect(x, sigma), gaussian_filter1d(x, sigma=sigma),
                                   rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", G["core"], ids=lambda c: c["name"])
def test_oracle_pipeline_matches_reference_golden(case):
    a = pause_audio(case["seed"], case["seconds"])
    assert [list(s) for s in segment_speech_oracle(a, 16000, **case["kwargs"])] == case["segments"]


def test_frame_count_and_filterbank_bins_match_oracle():
    from speechcatcher_b200 import _lib
    from speechcatcher_b200.simple_endpointing import num_frames
    for n in (1, 399, 400, 401, 559, 560, 561, 16000, 960001, 57600000):
        assert num_frames(n) == psf_n_frames(n), n
        if n <= 16000:
            assert num_frames(n) == psf_logfbank(np.ones(n, np.int16)).shape[0], n
    bins = (C.c_double * 28)()
    _lib.check(_lib.load().sc_segment_filterbank_bins(bins))
    assert list(bins) == psf_filterbank_bins().tolist()


def test_segmenter_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from speechcatcher_b200.simple_endpointing import segment_speech
    with pytest.raises(RuntimeError):
        segment_speech(np.zeros(16000 * 61, np.int16), 16000)
