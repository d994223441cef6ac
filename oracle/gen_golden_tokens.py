"""Generate tests/golden/tokens.json by running the REFERENCE on a model directory that holds a SentencePiece model
(build container only).  TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.gen_golden_tokens

tests/golden/bpe_unigram1024.model is a 1024-piece unigram model trained offline on a synthetic corpus (there is no
network for the real `bpe.model` files).  Pins the token-list construction (speech2text_streaming.py:97-124) and the
ids -> text assembly (:520-535) of the reference; the oracle must reproduce both before the file is written.
"""
from __future__ import annotations

import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle.speech2text import OracleSpeech2Text  # noqa: E402
from speechcatcher_b200.synthetic import make_model_dir, synth_audio  # noqa: E402

BPE = REPO / "tests" / "golden" / "bpe_unigram1024.model"


def make_dir_with_bpe(path, arch="m_d2", eos_bias=7.0):
    md = make_model_dir(path, arch, seed=0, eos_bias=eos_bias)
    shutil.copy(BPE, Path(md) / "bpe.model")
    return md


def main():
    sys.path.insert(0, "/root/reference")
    from speechcatcher.speech2text_streaming import Speech2TextStreaming as Ref
    out = {}
    with tempfile.TemporaryDirectory() as td:
        md = make_dir_with_bpe(Path(td) / "m")
        ref, orc = Ref(md, beam_size=5, device="cpu"), OracleSpeech2Text(md, beam_size=5)
        assert ref.token_list == orc.token_list and len(ref.token_list) == 1024
        out["token_list"] = ref.token_list
        audio = synth_audio(2, 10 * 8192 + 500)
        calls = []
        for i in range(0, len(audio), 8192):
            fin = i + 8192 >= len(audio)
            r = ref(audio[i:i + 8192].copy(), is_final=fin, finalize_all=fin)
            o = orc(audio[i:i + 8192].copy(), is_final=fin, finalize_all=fin)
            r = [(t, toks, [int(x) for x in ids]) for t, toks, ids in r]
            assert [tuple(x) for x in o] == r, (i, o[:1], r[:1])
            calls.append([list(x) for x in r])
        out["audio"] = dict(seed=2, n=len(audio))
        out["calls"] = calls
        print("final text:", repr(calls[-1][0][0][:100]))
    p = REPO / "tests" / "golden" / "tokens.json"
    p.write_text(json.dumps(out, ensure_ascii=False))
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
