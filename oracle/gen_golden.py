"""Generate golden vectors by running the REFERENCE itself (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/<case>.npz

For every case this drives /root/reference's `Speech2TextStreaming` (CPU, fp32)
chunk by chunk on a synthetic model dir + synthetic audio, records per call the
normalised features, the encoder output, the beam (yseq / fp64 score / xpos) and
the returned tuples, plus the first decode steps' decoder / CTC / combined score
rows, and checks that the oracle port reproduces all of it before writing.
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from speechcatcher_b200.synthetic import make_model_dir, synth_audio  # noqa: E402

CHUNK = 8192

# name: (arch, beam, n_samples, audio kind, pattern, use_bbd, sharpen, n_trace_steps)
CASES = {
    "m_d2_b5_6s": ("m_d2", 5, 6 * 16000, "noise", "A", False, 1.0, 6),
    "xl_d4_b10_cli": ("xl_d4", 10, 10 * CHUNK + 3000, "tones", "B", False, 1.0, 4),
    "m_d2_b5_bbd": ("m_d2", 5, 6 * 16000, "noise", "A", True, 1.0, 0),
    "xl_d4_b10_sharp": ("xl_d4", 10, 5 * 16000, "noise", "A", False, 8.0, 0),
    "m_d2_b5_short": ("m_d2", 5, 5000, "noise", "A", False, 1.0, 2),
    "xl_b10_4s": ("xl", 10, 4 * 16000, "noise", "A", False, 1.0, 0),
    "m_b5_8s": ("m", 5, 8 * 16000, "tones", "A", False, 1.0, 0),
    # EOS-biased weights (synthetic.make_state_dict eos_bias): hypotheses end regularly, so the <eos> branches of the
    # search (ended hypotheses, results of non-final / intermediate-final calls) carry data
    "xl_d4_b10_eos": ("xl_d4", 10, 6 * 16000 + 999, "noise", "A", False, 1.0, 2, 7.0),
    # live-server pattern (speechcatcher_server.py:252-270): utterances finalised mid-stream, NO reset in between
    "m_d2_b5_eos_live": ("m_d2", 5, 16 * CHUNK, "noise", "L", False, 1.0, 0, 7.0),
    # the benchmark's own regime: 60 s utterances that run into the global 500-step cap
    # (beam_search.py:701,821,827-836; T up to 1500, L up to 500) and a full-depth XL case; "lite" records hold the
    # beams of every call (int16 arrays) but no feature / encoder tensors
    "m_d2_b5_60s": ("m_d2", 5, 60 * 16000, "noise", "A", False, 1.0, 0, 0.0, True),
    "xl_d4_b10_60s": ("xl_d4", 10, 60 * 16000, "noise", "A", False, 1.0, 0, 0.0, True),
    "xl_b10_20s": ("xl", 10, 20 * 16000, "noise", "A", False, 1.0, 0, 0.0, True),
    # BASELINE.json configs[4] (beam 20: two hypothesis tiles in the tensor-core decoder attention) and configs[2]
    # (the L architecture, full depth: 18 encoder / 8 decoder layers)
    "xl_d4_b20_6s": ("xl_d4", 20, 6 * 16000 + 321, "noise", "A", False, 1.0, 2),
    "l_b10_6s": ("l", 10, 6 * 16000, "tones", "A", False, 1.0, 0),
}


def chunk_plan(n_samples: int, pattern: str):
    """[(start, end, is_final)] -- A: last real chunk is final; B: CLI order, a trailing
    empty final call (reference speechcatcher.py:436-446, 578-589)."""
    calls = []
    n_chunks = (n_samples + CHUNK - 1) // CHUNK
    for i in range(n_chunks):
        s, e = i * CHUNK, min((i + 1) * CHUNK, n_samples)
        calls.append((s, e, (pattern in ("A", "L") and i == n_chunks - 1) or (pattern == "L" and i in (5, 6, 11))))
    if pattern == "B":
        calls.append((n_samples, n_samples, True))
    return calls


def run_reference(model_dir, beam, audio, calls, use_bbd, n_trace):
    sys.path.insert(0, "/root/reference")
    from speechcatcher.speech2text_streaming import Speech2TextStreaming

    ref = Speech2TextStreaming(model_dir, beam_size=beam, ctc_weight=0.3, device="cpu", use_bbd=use_bbd)
    cap = {}
    orig_frontend = ref.apply_frontend

    def frontend_hook(*a, **k):
        out = orig_frontend(*a, **k)
        cap["feats"] = out[0]
        return out

    ref.apply_frontend = frontend_hook
    ref.beam_search.encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc", o[0]))
    bs = ref.beam_search.beam_search
    orig_score = bs.batch_score_hypotheses
    trace = []

    def score_hook(hyps, enc_out, pre_beam_size=40):
        comb, st, ind = orig_score(hyps, enc_out, pre_beam_size)
        if len(trace) < n_trace:
            trace.append(dict(dec=ind["decoder"].numpy().copy(), ctc=ind["ctc"].numpy().copy(),
                              comb=comb.numpy().copy(), T=enc_out.size(1)))
        return comb, st, ind

    bs.batch_score_hypotheses = score_hook
    rec = []
    t0 = time.perf_counter()
    for (s, e, fin) in calls:
        cap.clear()
        res = ref(audio[s:e], is_final=fin, finalize_all=fin)
        hyps = ref.beam_state.hypotheses if ref.beam_state is not None else []
        rec.append(dict(
            feats=None if cap.get("feats") is None else cap["feats"][0].numpy().copy(),
            enc=None if cap.get("enc") is None else cap["enc"][0].numpy().copy(),
            yseq=[h.yseq.tolist() for h in hyps], score=[float(h.score) for h in hyps],
            xpos=[h.xpos.tolist() for h in hyps], results=[list(map(int, r[2])) for r in res],
            process_idx=ref.beam_search.process_idx))
    return rec, trace, time.perf_counter() - t0


def run_oracle(model_dir, beam, audio, calls, use_bbd, n_trace):
    from oracle.speech2text import OracleSpeech2Text

    trace = []

    def tr(d):
        if len(trace) < n_trace:
            trace.append(dict(dec=d["dec"].numpy().copy(), ctc=d["ctc"].numpy().copy(),
                              comb=d["combined"].numpy().copy(), T=d["T"]))

    o = OracleSpeech2Text(model_dir, beam_size=beam, ctc_weight=0.3, use_bbd=use_bbd, trace=tr)
    rec = []
    t0 = time.perf_counter()
    for (s, e, fin) in calls:
        o.search.last_enc_out = None
        res = o(audio[s:e], is_final=fin, finalize_all=fin)
        hyps = o.hyps or []
        called = o.last_feats is not None
        enc = o.search.last_enc_out if called else None
        rec.append(dict(
            feats=None if o.last_feats is None else o.last_feats.numpy().copy(),
            enc=None if enc is None else enc[0].numpy().copy(),
            yseq=[list(h.yseq) for h in hyps], score=[float(h.score) for h in hyps],
            xpos=[list(h.xpos) for h in hyps], results=[list(r[2]) for r in res],
            process_idx=o.search.process_idx))
    return rec, trace, time.perf_counter() - t0


def compare(ref, orc, name):
    """Max deviations between two call records; raises on any n-best mismatch."""
    dev = dict(feats=0.0, enc=0.0, score=0.0)
    for i, (a, b) in enumerate(zip(ref, orc)):
        for k in ("feats", "enc"):
            if (a[k] is None) != (b[k] is None):
                # the reference skips the encoder when < 3 feature frames arrive
                if k == "enc" and (a[k] is None or a[k].shape[0] == 0) and (b[k] is None or b[k].shape[0] == 0):
                    continue
                raise AssertionError(f"{name} call {i}: {k} presence differs")
            if a[k] is not None:
                assert a[k].shape == b[k].shape, (name, i, k, a[k].shape, b[k].shape)
                if a[k].size:
                    dev[k] = max(dev[k], float(np.abs(a[k] - b[k]).max()))
        assert a["yseq"] == b["yseq"], f"{name} call {i}: yseq differs"
        assert a["xpos"] == b["xpos"], f"{name} call {i}: xpos differs"
        assert a["results"] == b["results"], f"{name} call {i}: results differ"
        assert a["process_idx"] == b["process_idx"], f"{name} call {i}: process_idx differs"
        if a["score"]:
            dev["score"] = max(dev["score"], float(np.abs(np.array(a["score"]) - np.array(b["score"])).max()))
    return dev


def pack(rec, trace, meta, lite=False):
    out = {"meta": np.array(json.dumps(meta))}
    for i, r in enumerate(rec):
        out[f"c{i}_score"] = np.array(r["score"], dtype=np.float64)
        out[f"c{i}_process_idx"] = np.array(r["process_idx"])
        if lite:
            # beams as int16 arrays [n_hyp][len] (block-synchronous: equal lengths); "called" = the frontend emitted
            out[f"c{i}_called"] = np.array(r["feats"] is not None)
            out[f"c{i}_nfeat"] = np.array(0 if r["feats"] is None else r["feats"].shape[0])
            out[f"c{i}_nenc"] = np.array(0 if r["enc"] is None else r["enc"].shape[0])
            out[f"c{i}_yseq"] = np.array(r["yseq"], dtype=np.int16).reshape(len(r["yseq"]), -1)
            out[f"c{i}_xpos"] = np.array(r["xpos"], dtype=np.int16).reshape(len(r["xpos"]), -1)
            out[f"c{i}_results"] = np.array(json.dumps(r["results"]))
            continue
        if r["feats"] is not None:
            out[f"c{i}_feats"] = r["feats"].astype(np.float32)
        if r["enc"] is not None:
            out[f"c{i}_enc"] = r["enc"].astype(np.float32)
        out[f"c{i}_json"] = np.array(json.dumps(dict(yseq=r["yseq"], xpos=r["xpos"], results=r["results"])))
    for j, t in enumerate(trace):
        for k in ("dec", "ctc", "comb"):
            out[f"t{j}_{k}"] = t[k].astype(np.float32)
        out[f"t{j}_T"] = np.array(t["T"])
    return out


def main(only=None):
    torch.manual_seed(0)
    outdir = REPO / "tests" / "golden"
    outdir.mkdir(parents=True, exist_ok=True)
    for name, case in CASES.items():
        arch, beam, n, kind, pattern, bbd, sharpen, n_trace = case[:8]
        eos_bias = case[8] if len(case) > 8 else 0.0
        lite = bool(case[9]) if len(case) > 9 else False
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as td:
            md = make_model_dir(td, arch, seed=0, sharpen=sharpen, eos_bias=eos_bias)
            audio = synth_audio(0, n, kind)
            calls = chunk_plan(n, pattern)
            ref, rtrace, tr = run_reference(md, beam, audio, calls, bbd, n_trace)
            orc, otrace, to = run_oracle(md, beam, audio, calls, bbd, n_trace)
        dev = compare(ref, orc, name)
        for a, b in zip(rtrace, otrace):
            for k in ("dec", "ctc", "comb"):
                d = float(np.abs(a[k] - b[k]).max())
                dev["trace_" + k] = max(dev.get("trace_" + k, 0.0), d)
        meta = dict(arch=arch, beam=beam, n_samples=n, kind=kind, pattern=pattern, use_bbd=bbd,
                    sharpen=sharpen, eos_bias=eos_bias, seed=0, stream=0, calls=calls, chunk=CHUNK,
                    ref_seconds=tr, oracle_seconds=to, oracle_vs_ref_dev=dev, lite=lite,
                    final_process_idx=ref[-1]["process_idx"], torch=torch.__version__)
        np.savez_compressed(outdir / f"{name}.npz", **pack(ref, rtrace, meta, lite))
        print(f"{name}: ref {tr:.1f}s oracle {to:.1f}s calls {len(calls)} "
              f"final_len {len(ref[-1]['yseq'][0]) if ref[-1]['yseq'] else 0} dev {dev}")


if __name__ == "__main__":
    main(sys.argv[1:] or None)
