"""GPU: direct parity of the per-step score tensors (SURVEY.md rows K6 CTC head, K7 CTC prefix scorer, K8 decoder step,
K9 combine + pre-beam) with the rows the reference's own `batch_score_hypotheses` produced (beam_search.py:71-185),
recorded in the goldens as t<j>_dec / t<j>_ctc / t<j>_comb by oracle/gen_golden.py.

Tolerance (BASELINE.json north_star, fp32 modes): decoder log-probs and CTC scores 1e-3 absolute; the pre-beam ids and
their order are exact."""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, load_golden, model_dir

pytestmark = pytest.mark.gpu

TRACED = [c for c in GOLDEN_CASES if load_golden(c)[2]]
LOGZERO = -1e10


def _trace_case(case, dtype):
    from speechcatcher_b200 import Speech2TextStreaming
    from speechcatcher_b200.synthetic import synth_audio
    meta, calls, trace = load_golden(case)
    md = model_dir(meta["arch"], meta["seed"], meta["sharpen"], meta.get("eos_bias", 0.0))
    audio = synth_audio(meta["stream"], meta["n_samples"], meta["kind"])
    max_chunk = max(8192, max(e - s for s, e, _ in meta["calls"]))
    gpu = Speech2TextStreaming(md, beam_size=meta["beam"], ctc_weight=0.3, device="cuda:0", use_bbd=meta["use_bbd"],
                               max_chunk=max_chunk, dtype=dtype)
    gpu.group.trace_begin(len(trace))
    for (s, e, fin) in meta["calls"]:
        gpu(audio[s:e], is_final=fin, finalize_all=fin)
    got = gpu.group.trace_end()
    return meta, trace, got


@pytest.mark.parametrize("dtype", ["float32_simt", "float32_tc"])
@pytest.mark.parametrize("case", TRACED)
def test_step_scores_match_reference_rows(case, dtype):
    meta, trace, got = _trace_case(case, dtype)
    assert len(got) == len(trace), f"{len(got)} traced iterations, golden holds {len(trace)}"
    w_dec, w_ctc = np.float32(0.7), np.float32(0.3)
    worst = dict(dec=0.0, ctc=0.0, comb=0.0)
    for j, (g, t) in enumerate(zip(got, trace)):
        n = t["dec"].shape[0]
        assert len(g["rows"]) == n, f"step {j}: {len(g['rows'])} active rows vs {n} hypotheses in the reference"
        assert [h for _, h in g["rows"]] == list(range(n))
        assert int(g["Tb"][0]) == t["T"], f"step {j}: memory length {int(g['Tb'][0])} vs {t['T']}"
        # K8: decoder log-probs over the whole vocabulary
        d = np.abs(g["logp"] - t["dec"]).max()
        worst["dec"] = max(worst["dec"], float(d))
        assert d <= 1e-3, f"step {j}: decoder log-probs differ by {d}"
        # K9: pre-beam = top-40 of w_dec * dec, same ids in the same order
        want_ids = np.argsort(-(w_dec * t["dec"]), axis=1, kind="stable")[:, :40]
        assert (g["pre_ids"] == want_ids).all(), f"step {j}: pre-beam ids differ"
        # K6 + K7: CTC prefix scores of the 40 candidates (+ <eos>); blank is forced to logzero
        eos = t["dec"].shape[1] - 1
        for r in range(n):
            ids = g["pre_ids"][r]
            ctc = g["psi"][r] - g["s_prev"][r]
            ctc = np.where(ids == 0, np.float32(LOGZERO) - g["s_prev"][r], ctc)
            ctc = np.where(ids == eos, g["psi_eos"][r] - g["s_prev"][r], ctc)
            ref = t["ctc"][r, ids]
            big = np.abs(ref) > 1e9                   # logzero entries: compare loosely (fp32 spacing at 1e10 is 1024)
            assert (np.abs(ctc[big] - ref[big]) <= 2048).all(), f"step {j} row {r}: logzero entries"
            e = np.abs(ctc[~big] - ref[~big]).max() if (~big).any() else 0.0
            worst["ctc"] = max(worst["ctc"], float(e))
            assert e <= 1e-3, f"step {j} row {r}: CTC prefix scores differ by {e}"
            e_eos = abs(float(g["psi_eos"][r] - g["s_prev"][r]) - float(t["ctc"][r, eos]))
            assert e_eos <= 1e-3, f"step {j} row {r}: CTC <eos> score differs by {e_eos}"
            comb = w_dec * g["logp"][r, ids] + w_ctc * ctc
            ec = np.abs(comb[~big] - t["comb"][r, ids][~big]).max() if (~big).any() else 0.0
            worst["comb"] = max(worst["comb"], float(ec))
            assert ec <= 1e-3, f"step {j} row {r}: combined scores differ by {ec}"
    print(f"\n[step-trace {case} {dtype}] max |dec| {worst['dec']:.2e}  |ctc| {worst['ctc']:.2e}  |comb| {worst['comb']:.2e}")
