"""GPU: a model directory with a SentencePiece model through the drop-in facade: texts, token
strings and ids equal the reference-generated golden."""
import json
import shutil
from pathlib import Path

import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
G = json.loads((GOLDEN / "tokens.json").read_text())


def test_load_model_with_tokenizer_matches_reference_golden(tmp_path):
    from speechcatcher_b200.model_files import load_model
    from speechcatcher_b200.synthetic import make_model_dir, synth_audio
    d = make_model_dir(tmp_path / "unpacked" / "exp" / "asr_train", "m_d2", seed=0, eos_bias=7.0)
    shutil.copy(GOLDEN / "bpe_unigram1024.model", Path(d) / "bpe.model")
    s2t = load_model(str(tmp_path / "unpacked"), device="cuda", beam_size=5, quiet=True)
    assert s2t.token_list == G["token_list"] and s2t.mean is not None
    audio = synth_audio(G["audio"]["seed"], G["audio"]["n"])
    for k, i in enumerate(range(0, len(audio), 8192)):
        fin = i + 8192 >= len(audio)
        got = s2t(audio[i:i + 8192], is_final=fin, finalize_all=fin)
        assert [[r[0], r[1], r[2]] for r in got] == G["calls"][k], k
    text, toks, ids, pos, hyp = got[0]
    assert "▁" not in text and len(pos) == len(ids)
