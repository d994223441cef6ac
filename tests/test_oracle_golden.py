"""CPU: the oracle port replayed against golden vectors produced by the reference itself
(oracle/gen_golden.py).  Tolerances: BASELINE.json north_star -- features 1e-4 relative,
log-probs 1e-3, n-best tokens / timestamps / order exact."""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, load_golden, model_dir

# full-depth / long cases the generator already checked against the reference; replayed on the GPU only
SLOW = ("xl_b10_4s", "m_b5_8s", "xl_b10_20s", "xl_d4_b10_60s")
FAST = [c for c in GOLDEN_CASES if c not in SLOW]


@pytest.mark.parametrize("case", FAST + ["xl_b10_4s"])
def test_oracle_replays_reference(case):
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200.synthetic import synth_audio

    meta, calls, trace = load_golden(case)
    md = model_dir(meta["arch"], meta["seed"], meta["sharpen"], meta.get("eos_bias", 0.0))
    audio = synth_audio(meta["stream"], meta["n_samples"], meta["kind"])
    got_trace = []
    o = OracleSpeech2Text(md, beam_size=meta["beam"], ctc_weight=0.3, use_bbd=meta["use_bbd"],
                          trace=lambda d: got_trace.append(d) if len(got_trace) < len(trace) else None)
    for (s, e, fin), g in zip(meta["calls"], calls):
        o.search.last_enc_out = None
        res = o(audio[s:e], is_final=fin, finalize_all=fin)
        if meta.get("lite"):                       # long-utterance goldens: shapes and beams only
            assert (o.last_feats is not None) == g["called"]
            if g["called"]:
                assert o.last_feats.shape[0] == g["n_feat"]
        elif g["feats"] is not None:
            f = o.last_feats.numpy()
            assert f.shape == g["feats"].shape
            assert np.abs(f - g["feats"]).max() <= 1e-4 * max(1.0, np.abs(g["feats"]).max())
        else:
            assert o.last_feats is None
        if g["enc"] is not None and g["enc"].size:
            e_ = o.search.last_enc_out[0].numpy()
            assert e_.shape == g["enc"].shape
            np.testing.assert_allclose(e_, g["enc"], atol=1e-3, rtol=0)
        hyps = o.hyps or []
        assert [list(h.yseq) for h in hyps] == g["yseq"]
        assert [list(h.xpos) for h in hyps] == g["xpos"]
        np.testing.assert_allclose([h.score for h in hyps], g["score"], atol=1e-3, rtol=1e-6)
        assert [list(r[2]) for r in res] == g["results"]
        assert o.search.process_idx == g["process_idx"]
    for a, b in zip(got_trace, trace):
        real = b["ctc"] > -1e9
        np.testing.assert_allclose(a["dec"].numpy(), b["dec"], atol=1e-3, rtol=0)
        np.testing.assert_allclose(a["ctc"].numpy()[real], b["ctc"][real], atol=1e-3, rtol=1e-5)


def test_long_goldens_reach_the_step_cap():
    """The 60 s goldens end at the reference's global step cap (beam_search.py:701,821: process_idx < 500, rewound to
    499 at the end of every block), the regime every stream of the benchmark runs in."""
    for case in ("m_d2_b5_60s", "xl_d4_b10_60s"):
        meta, calls, _ = load_golden(case)
        assert meta["final_process_idx"] == 499 and calls[-1]["process_idx"] == 499, case
        assert len(calls[-1]["yseq"][0]) > 500, case


def test_mel_filterbank_matches_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    from oracle.frontend import hann_window, mel_filterbank
    ref = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney")
    assert torch.equal(mel_filterbank(), ref)
    assert torch.allclose(hann_window(), torch.hann_window(400), atol=1e-7)
