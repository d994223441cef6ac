"""CPU: pre-computed feature input (speech2text_streaming.py:438-450) -- the oracle against goldens produced by the
reference (oracle/gen_golden_feats.py)."""
import json

import numpy as np
import pytest

from helpers import GOLDEN, model_dir
from oracle.gen_golden_feats import feature_chunks

G = json.loads((GOLDEN / "feats_input.json").read_text())


@pytest.mark.parametrize("case", G, ids=lambda c: c["name"])
def test_oracle_feature_input_matches_reference_golden(case):
    from oracle.speech2text import OracleSpeech2Text
    orc = OracleSpeech2Text(model_dir(case["arch"], eos_bias=case["eos_bias"]), beam_size=case["beam"])
    for f, fin, g in zip(feature_chunks(case["seed"], case["frames"]), case["finals"], case["calls"]):
        res = orc(f[None] if case["batched"] else f, is_final=fin, finalize_all=fin)
        assert [list(h.yseq) for h in orc.hyps] == g["yseq"]
        assert [list(h.xpos) for h in orc.hyps] == g["xpos"]
        np.testing.assert_allclose([h.score for h in orc.hyps], g["score"], atol=1e-3, rtol=0)
        assert [list(r[2]) for r in res] == g["results"] and orc.search.process_idx == g["process_idx"]
