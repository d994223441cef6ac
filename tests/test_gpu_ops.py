"""GPU: the per-operator C-ABI entry points against the oracle's restatement of the same reference functions
(SURVEY.md 8(b) per-op list): K1 sc_frontend_fbank_mvn, K7 sc_ctc_prefix_step.

Tolerances (BASELINE.json north_star): features 1e-4 relative; CTC prefix scores 1e-3 (log domain, fp32)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(a, device="cuda")


@pytest.mark.parametrize("n_samples,with_stats", [(8192, True), (16000 * 3 + 77, True), (401, False), (5000, False)])
def test_frontend_fbank_mvn_matches_oracle(n_samples, with_stats):
    from oracle.frontend import hann_window, log_mel, mel_filterbank
    from speechcatcher_b200 import _lib
    from speechcatcher_b200.synthetic import make_feats_stats, synth_audio
    lib = _lib.load()
    window, mel = hann_window(), mel_filterbank()
    wave = torch.from_numpy(synth_audio(3, n_samples, "tones"))
    want = log_mel(wave, window, mel)
    mean = std = None
    if with_stats:
        st = make_feats_stats()
        mean = st["sum"] / st["count"]
        std = np.sqrt(np.maximum(st["sum_square"] / st["count"] - mean ** 2, 1e-10))
        want = torch.from_numpy(((want.numpy() - mean) / std)).to(torch.float32)     # numpy fp64 round trip (:355-358)
    ws = torch.zeros(lib.sc_frontend_workspace_bytes(), dtype=torch.uint8, device="cuda")
    w_h, m_h = window.contiguous(), mel.contiguous()
    _lib.check(lib.sc_frontend_init(ws.data_ptr(), w_h.data_ptr(), m_h.data_ptr()), "frontend_init")
    wd = _dev(wave)
    n_frames = 1 + n_samples // 160
    feats = torch.full((n_frames, 80), float("nan"), device="cuda")
    md, sd = (_dev(mean), _dev(std)) if with_stats else (None, None)
    got_n = C.c_int32()
    _lib.check(lib.sc_frontend_fbank_mvn(ws.data_ptr(), wd.data_ptr(), n_samples, md.data_ptr() if with_stats else None,
                                         sd.data_ptr() if with_stats else None, feats.data_ptr(), C.byref(got_n), None),
               "frontend_fbank_mvn")
    torch.cuda.synchronize()
    assert got_n.value == n_frames == want.shape[0]
    err = (feats.cpu() - want).abs().max().item()
    assert err <= 1e-4 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("T,L,n_hyp,seed", [(24, 0, 1, 0), (24, 3, 5, 1), (200, 57, 10, 2), (751, 240, 10, 3), (40, 60, 10, 4)])
def test_ctc_prefix_step_matches_oracle(T, L, n_hyp, seed):
    """One CTCPrefixScoreTH.__call__ (ctc_prefix_score_full.py:88-291) on random emissions and forward variables:
    log_psi of the 40 candidates, the <eos> entry and the per-candidate forward variables r (T, 2)."""
    from oracle.ctc_prefix import LOGZERO, CTCPrefixOracle
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    V, K = 1024, 40
    g = torch.Generator().manual_seed(seed)
    orc = CTCPrefixOracle({"ctc.ctc_lo.weight": torch.zeros(V, 4), "ctc.ctc_lo.bias": torch.zeros(V)})
    # rows of the first block are log-softmax, later rows raw logits (quirk Q1): mix both kinds
    x = torch.randn(T, V, generator=g) * 2.0
    x[:24] = torch.log_softmax(x[:24], dim=-1)
    orc.x, orc.T = x, T
    yseqs = torch.randint(2, V - 1, (n_hyp, L + 1), generator=g)
    yseqs[:, 0] = V - 1
    ids = torch.stack([torch.randperm(V - 2, generator=g)[:K] + 1 for _ in range(n_hyp)])
    ids[:, 0] = yseqs[:, -1] if L > 0 else ids[:, 0]            # the repeated-token branch (phi = r_prev blank only)
    if L == 0:
        states = [None] * n_hyp
        r_prev = torch.full((n_hyp, T, 2), LOGZERO)
        r_prev[:, :, 1] = torch.cumsum(x[:, 0], 0).unsqueeze(0)
    else:
        # plausible forward variables: decreasing log-probabilities with some logzero entries at the start
        r_prev = -torch.rand(n_hyp, T, 2, generator=g) * 30.0 - torch.arange(T).view(1, T, 1) * 0.5
        r_prev[:, : min(L, T) - 1, :] = LOGZERO
        states = [(r_prev[h].clone(), torch.zeros(V)) for h in range(n_hyp)]
    scores, (r_ref, log_psi, idmap) = orc.score_partial(yseqs, ids, states)
    psi = torch.full((n_hyp, K), float("nan"), device="cuda")
    psi_eos = torch.full((n_hyp,), float("nan"), device="cuda")
    r_new = torch.full((n_hyp, K, T, 2), float("nan"), device="cuda")
    xd, rd = _dev(x.contiguous()), _dev(r_prev.contiguous())
    last = _dev(yseqs[:, -1].to(torch.int32).contiguous())
    idd = _dev(ids.to(torch.int32).contiguous())
    _lib.check(lib.sc_ctc_prefix_step(xd.data_ptr(), T, V, rd.data_ptr(), last.data_ptr(), L, idd.data_ptr(), n_hyp,
                                      psi.data_ptr(), psi_eos.data_ptr(), r_new.data_ptr(), None), "ctc_prefix_step")
    torch.cuda.synchronize()
    want_psi = torch.gather(log_psi, 1, ids)
    real = want_psi > -1e9
    assert torch.isfinite(psi).all()
    if real.any():                                                             # T <= L: every prefix score is logzero
        assert (psi.cpu()[real] - want_psi[real]).abs().max().item() <= 1e-3
    assert ((psi.cpu()[~real] - want_psi[~real]).abs() <= 2048).all()          # logzero entries (fp32 spacing at 1e10)
    assert (psi_eos.cpu() - log_psi[:, V - 1]).abs().max().item() <= 1e-3
    # forward variables: the reference's r is (T, 2, n_hyp, K)
    want_r = r_ref.permute(2, 3, 0, 1)
    got_r = r_new.cpu()
    big = want_r < -1e9
    if (~big).any():
        assert (got_r[~big] - want_r[~big]).abs().max().item() <= 1e-3
    assert (got_r[big] < -1e9).all()
