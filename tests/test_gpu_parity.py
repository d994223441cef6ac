"""GPU: the CUDA path (through the C ABI) against the reference-generated goldens and the live oracle.

Tolerances (BASELINE.json north_star, fp32 mode): features 1e-4 relative, encoder outputs / log-probs
1e-3 absolute, n-best token sequences, token timestamps (xpos) and beam order exact."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, load_golden, model_dir

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.as_tensor(a, device="cuda")


def test_layernorm_and_linear_ops():
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(333, 256, generator=g)
    w, b = torch.randn(256, generator=g), torch.randn(256, generator=g)
    xd, wd, bd = _t(x), _t(w), _t(b)
    y = torch.empty_like(xd)
    _lib.check(lib.sc_layernorm_f32(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), 333, 256, None))
    ref = torch.nn.functional.layer_norm(x, (256,), w, b, 1e-12)
    torch.cuda.synchronize()
    assert (y.cpu() - ref).abs().max() < 1e-5
    for (m, n, k, relu, res) in [(333, 768, 256, 0, False), (77, 256, 2048, 0, True), (130, 2048, 256, 1, False),
                                 (5, 1024, 256, 0, False), (64, 256, 4864, 0, False)]:
        a = torch.randn(m, k, generator=g)
        W = torch.randn(n, k, generator=g) / k ** 0.5
        bias = torch.randn(n, generator=g)
        r = torch.randn(m, n, generator=g)
        ad, Wd, biasd, rd = _t(a), _t(W), _t(bias), _t(r)
        out = rd.clone() if res else torch.empty(m, n, device="cuda")
        _lib.check(lib.sc_linear_f32(ad.data_ptr(), Wd.data_ptr(), biasd.data_ptr(), out.data_ptr() if res else None,
                                     out.data_ptr(), m, n, k, relu, None))
        want = torch.nn.functional.linear(a.double(), W.double(), bias.double())
        if relu:
            want = want.relu()
        if res:
            want = want + r.double()
        torch.cuda.synchronize()
        assert (out.cpu().double() - want).abs().max() < 2e-5, (m, n, k)


def _run_case(case, check_internals=True, dtype="float32"):
    from speechcatcher_b200 import Speech2TextStreaming
    from speechcatcher_b200.synthetic import synth_audio
    meta, calls, _ = load_golden(case)
    md = model_dir(meta["arch"], meta["seed"], meta["sharpen"], meta.get("eos_bias", 0.0))
    audio = synth_audio(meta["stream"], meta["n_samples"], meta["kind"])
    max_chunk = max(8192, max(e - s for s, e, _ in meta["calls"]))
    gpu = Speech2TextStreaming(md, beam_size=meta["beam"], ctc_weight=0.3, device="cuda:0", use_bbd=meta["use_bbd"],
                               max_chunk=max_chunk, dtype=dtype)
    grp = gpu.group
    cap_feat = grp.buffer("featbuf").numel() // 80
    enc_seen = 0
    for ci, ((s, e, fin), g) in enumerate(zip(meta["calls"], calls)):
        res = gpu(audio[s:e], is_final=fin, finalize_all=fin)
        plan = grp.last_plan(0)
        lite = bool(meta.get("lite"))              # long-utterance goldens hold beams and shapes only
        called = g["called"] if lite else g["feats"] is not None
        assert bool(plan.called) == called, f"call {ci}"
        if not called:
            assert res == []
            continue
        assert plan.n_feat == (g["n_feat"] if lite else g["feats"].shape[0]), f"call {ci}"
        n_enc = g["n_enc"] if lite else (0 if g["enc"] is None else g["enc"].shape[0])
        assert plan.n_enc_out == n_enc, f"call {ci}: enc frames {plan.n_enc_out} != {n_enc}"
        if check_internals and n_enc and not lite:
            enc = grp.buffer("encbuf").view(-1, 256)[enc_seen: enc_seen + n_enc].cpu().numpy()
            np.testing.assert_allclose(enc, g["enc"], atol=1e-3, rtol=0, err_msg=f"call {ci} encoder output")
        enc_seen += n_enc
        ys, sc, xp, pidx = gpu.beam_state
        assert ys == g["yseq"], f"call {ci}: n-best token sequences differ"
        assert xp == g["xpos"], f"call {ci}: token timestamps differ"
        # scores are sums of up to 600 fp32 increments of magnitude ~7: 2e-3 absolute up to ~130 tokens, relative beyond
        np.testing.assert_allclose(sc, g["score"], atol=2e-3, rtol=2e-6, err_msg=f"call {ci} scores")
        assert pidx == g["process_idx"], f"call {ci}"
        assert [r[2] for r in res] == g["results"], f"call {ci}"
    return gpu


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES])
def test_golden_nbest_exact(case):
    _run_case(case, dtype="float32_simt")


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES])
def test_golden_nbest_exact_tensor_core(case):
    """Same bar (n-best, timestamps, order, process_idx exact; encoder 1e-3) with every Linear on the tcgen05 tensor
    cores as a split-fp16 GEMM (precision 2, csrc/kernels_gemm_x3.cu)."""
    _run_case(case, dtype="float32_tc")


def test_frontend_features_vs_golden():
    """Frontend features of the first calls: 1e-4 relative (the feature buffer holds carry + new frames)."""
    from speechcatcher_b200 import Speech2TextStreaming
    from speechcatcher_b200.synthetic import synth_audio
    meta, calls, _ = load_golden("xl_d4_b10_cli")
    md = model_dir(meta["arch"], meta["seed"], meta["sharpen"], meta.get("eos_bias", 0.0))
    audio = synth_audio(meta["stream"], meta["n_samples"], meta["kind"])
    gpu = Speech2TextStreaming(md, beam_size=meta["beam"], device="cuda:0")
    carry = 0
    for ci, ((s, e, fin), g) in enumerate(zip(meta["calls"], calls)):
        if e - s == 0 or g["feats"] is None:
            break
        gpu(audio[s:e], is_final=fin, finalize_all=fin)
        # frames were written at row `carry` before the carry move; after the move the buffer front holds
        # the last n_res frames, so compare against the tail of (previous carry + new frames)
        feats = gpu.group.buffer("featbuf").view(-1, 80).cpu().numpy()
        T = carry + g["feats"].shape[0]
        n_samples = T // 4 - 1
        if n_samples < 2:
            got = feats[carry:T]
            want = g["feats"]
            carry = T
        else:
            n_res = T % 4 + 8
            got = feats[:n_res]
            want = g["feats"][-n_res:] if n_res <= g["feats"].shape[0] else None
            carry = n_res
        if want is not None:
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got - want).max() <= 1e-4 * scale, f"call {ci}"
