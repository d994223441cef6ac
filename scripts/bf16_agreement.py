"""Measure how the bf16 tensor-core mode tracks the fp32 oracle (not a pytest file; prints a table)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helpers import GOLDEN_CASES, load_golden, model_dir


def main():
    from speechcatcher_b200 import Speech2TextStreaming
    from speechcatcher_b200.synthetic import synth_audio
    for case in GOLDEN_CASES:
        meta, calls, _ = load_golden(case)
        md = model_dir(meta["arch"], meta["seed"], meta["sharpen"], meta.get("eos_bias", 0.0))
        audio = synth_audio(meta["stream"], meta["n_samples"], meta["kind"])
        mc = max(8192, max(e - s for s, e, _ in meta["calls"]))
        gpu = Speech2TextStreaming(md, beam_size=meta["beam"], device="cuda:0", dtype="bfloat16", use_bbd=meta["use_bbd"], max_chunk=mc)
        enc_seen, enc_err, first_div = 0, 0.0, None
        for ci, ((s, e, fin), g) in enumerate(zip(meta["calls"], calls)):
            gpu(audio[s:e], is_final=fin, finalize_all=fin)
            if g["feats"] is None:
                continue
            n_enc = 0 if g["enc"] is None else g["enc"].shape[0]
            if n_enc:
                enc = gpu.group.buffer("encbuf").view(-1, 256)[enc_seen: enc_seen + n_enc].cpu().numpy()
                enc_err = max(enc_err, float(np.abs(enc - g["enc"]).max()))
            enc_seen += n_enc
            ys = gpu.beam_state[0]
            if first_div is None and ys and g["yseq"] and ys[0] != g["yseq"][0]:
                first_div = ci
        ys = gpu.beam_state[0]
        a, b = ys[0], calls[-1]["yseq"][0]
        common = 0
        for x, y in zip(a, b):
            if x != y:
                break
            common += 1
        print(f"{case:18s} enc_max_err {enc_err:.4f} best_equal {a == b} common_prefix {common}/{len(b)} first_div_call {first_div}")


if __name__ == "__main__":
    main()
