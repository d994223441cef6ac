"""A/B check of the bf16 mode's two decoder-attention implementations (SIMT vs mma.sync tensor cores).
Not a pytest file: prints the agreement of decoder log-probs and beams on a short multi-stream run."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from helpers import model_dir


def run(kind, arch, beam, n_streams, seconds):
    os.environ["SCB_ATTN"] = kind
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir(arch)
    n = int(seconds * 16000)
    audio = np.stack([synth_audio(50 + s, n) for s in range(n_streams)])
    grp = StreamGroup(md, n_streams=n_streams, beam_size=beam, device="cuda:0", dtype="bfloat16", max_seconds=seconds + 2)
    ids = np.arange(n_streams, dtype=np.int32)
    logs, beams, encs = [], [], []
    for i in range(0, n, 8192):
        fin = i + 8192 >= n
        ch = [audio[s, i:i + 8192] for s in range(n_streams)]
        grp.push(ids, ch, [fin] * n_streams)
        logs.append(grp.buffer("dlogp").view(-1, 1024)[: n_streams * beam].clone().cpu())
        beams.append([grp.beam(s)[0] for s in range(n_streams)])
    encs = grp.buffer("encbuf").view(n_streams, -1, 256)[:, : int(seconds * 25) - 8].clone().cpu()
    return logs, beams, encs


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    for arch, beam in (("xl_d4", 10), ("m_d2", 5)):
        a_logs, a_beams, a_enc = run("simt", arch, beam, 4, seconds)
        b_logs, b_beams, b_enc = run("mma", arch, beam, 4, seconds)
        print(f"{arch}: encoder output max |simt - mma| = {float((a_enc - b_enc).abs().max()):.5f} (scale {float(a_enc.abs().max()):.2f})")
        first_div = None
        for i, (x, y) in enumerate(zip(a_beams, b_beams)):
            if x != y and first_div is None:
                first_div = i
        # log-probs are comparable only while the beams agree
        upto = first_div if first_div is not None else len(a_logs)
        diffs = [float((a_logs[i] - b_logs[i]).abs().max()) for i in range(upto) if a_logs[i].abs().sum() > 0]
        same_final = sum(x[0] == y[0] for x, y in zip(a_beams[-1], b_beams[-1]))
        print(f"{arch} beam {beam}: pushes {len(a_logs)} first beam divergence at push {first_div}; "
              f"max |dlogp| while beams agree {max(diffs) if diffs else None}; same final 1-best {same_final}/4")


if __name__ == "__main__":
    main()
