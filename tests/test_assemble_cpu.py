"""CPU: the numpy result assembly (StreamGroup._assemble) against its line-by-line restatement of
speech2text_streaming.py:466-539, on random beams that contain every special case (blank / <unk> / <eos> inside and at
the end, empty hypotheses, ended and running hypotheses, with and without a token list)."""
import json
import time

import numpy as np
import pytest

from helpers import GOLDEN
from speechcatcher_b200.stream_group import StreamGroup


def _random_beam(rng, n, L):
    y = rng.integers(0, 1024, size=(n, L)).astype(np.int32)
    y[:, 0] = 1023
    y[rng.random((n, L)) < 0.1] = rng.choice([0, 1, 1023])
    ended = rng.random(n) < 0.5
    y[ended, -1] = 1023
    y[~ended & (y[:, -1] == 1023), -1] = 7
    xp = np.sort(rng.integers(0, 1500, size=(n, L)), axis=1).astype(np.int32)
    return y, rng.standard_normal(n).tolist(), xp, 3


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:4] == y[:4]
        assert x[4]["yseq"].tolist() == y[4]["yseq"].tolist() and x[4]["xpos"].tolist() == y[4]["xpos"].tolist()
        assert x[4]["score"] == y[4]["score"]


@pytest.mark.parametrize("with_tokens", [False, True])
def test_assembly_equals_rowwise(with_tokens):
    token_list = json.loads((GOLDEN / "tokens.json").read_text())["token_list"] if with_tokens else None
    rng = np.random.default_rng(0)
    for trial in range(60):
        n, L = int(rng.integers(1, 21)), int(rng.integers(1, 40))
        beam = _random_beam(rng, n, L)
        for is_final in (False, True):
            for finalize_all in (False, True):
                _same(StreamGroup._assemble(beam, is_final, finalize_all, token_list),
                      StreamGroup._assemble_rowwise(beam, is_final, finalize_all, token_list))
    assert StreamGroup._assemble(([], [], [], 0), True, True, token_list) == []
    # list-of-lists input (what StreamGroup.beam() returns) is accepted too
    y, sc, xp, p = _random_beam(rng, 5, 12)
    _same(StreamGroup._assemble((y.tolist(), sc, xp.tolist(), p), True, True, token_list),
          StreamGroup._assemble_rowwise((y, sc, xp, p), True, True, token_list))


def test_assembly_of_a_full_batch_is_fast():
    """256 streams x 10 hypotheses x 520 tokens (the bench's end-of-pass read-back) well under a second."""
    rng = np.random.default_rng(1)
    beams = [_random_beam(rng, 10, 520) for _ in range(64)]
    t = time.perf_counter()
    for b in beams:
        StreamGroup._assemble(b, True, True, None)
    assert (time.perf_counter() - t) * 4 < 2.0
