"""CPU: the C++ host planner (shape schedule of frontend / encoder / decode trigger) against the
oracle's actual tensor shapes on ragged chunk patterns, through the C ABI (no GPU calls)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import model_dir


def _oracle_shapes(chunks, finals, W):
    """Drive oracle frontend + encoder and mirror process_block's triggers (beam_search.py:551-634)."""
    from oracle.encoder import EncoderOracle
    from oracle.frontend import FrontendOracle
    fe = FrontendOracle(None, None)
    enc = EncoderOracle(W, 2, 4)
    enc_len, processed_block = 0, 0
    out = []
    for n, fin in zip(chunks, finals):
        feats = fe(torch.zeros(n), fin)
        if feats is None:
            out.append(dict(called=0, n_feat=0, n_enc_out=0, enc_len=enc_len, n_decode=0, last_T=0))
            continue
        n_feat = feats.size(0)
        n_enc = 0
        if n_feat >= 3:
            y = enc(feats.unsqueeze(0), fin)
            n_enc = y.size(1)
        enc_len += n_enc
        nd, last_T = 0, 0
        while enc_len > 0 and 24 + 16 * processed_block < enc_len:
            last_T = 24 + 16 * processed_block
            nd += 1
            processed_block += 1
        if fin and enc_len > 0:
            nd += 1
            last_T = enc_len
        out.append(dict(called=1, n_feat=n_feat, n_enc_out=n_enc, enc_len=enc_len, n_decode=nd, last_T=last_T))
    return out


@pytest.mark.parametrize("seed", range(16))
def test_planner_matches_oracle_shapes(seed):
    from oracle.speech2text import load_model_dir
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    W, _, _, _ = load_model_dir(model_dir("m_d2"))
    rng = np.random.default_rng(seed)
    n_calls = int(rng.integers(3, 14))
    kinds = [8192, 8192, 8192, 4000, 1600, 300, 100, 16000, 25600, 0]
    chunks = [int(rng.choice(kinds)) if rng.random() < 0.8 else int(rng.integers(1, 12000)) for _ in range(n_calls)]
    finals = [False] * n_calls
    finals[-1] = True
    if seed % 2 == 1:                                   # CLI pattern: trailing empty final call
        chunks[-1] = 0
    if seed >= 6:
        # live-server pattern (speechcatcher_server.py:252-270): utterances are finalised in the middle of a stream
        # and the stream simply continues WITHOUT reset -- frontend/encoder state restart, enc_buf keeps growing
        n_calls = int(rng.integers(10, 30))
        chunks = [int(rng.choice(kinds[:4])) for _ in range(n_calls)]
        finals = [bool(rng.random() < 0.2) for _ in range(n_calls)]
    try:
        want = _oracle_shapes(chunks, finals, W)
    except Exception:
        pytest.skip("the reference algorithm itself cannot process this chunk pattern")
    p = C.c_void_p()
    assert lib.sc_planner_create(1, C.byref(p)) == 0
    try:
        for i, (n, fin) in enumerate(zip(chunks, finals)):
            pl = _lib.ScStreamPlan()
            rc = lib.sc_planner_push(p, 0, n, int(fin), C.byref(pl))
            assert rc == 0, lib.sc_last_error()
            w = want[i]
            got = dict(called=pl.called, n_feat=pl.n_feat, n_enc_out=pl.n_enc_out, enc_len=pl.enc_len,
                       n_decode=pl.n_decode_blocks, last_T=pl.last_T)
            assert got == w, f"call {i} chunks={chunks}: {got} != {w}"
    finally:
        lib.sc_planner_destroy(p)


def test_planner_reset_and_8192_schedule():
    """49-50 feature frames per 8192-sample chunk, first encoder output on the 4th call,
    first decode block on the 5th."""
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    p = C.c_void_p()
    assert lib.sc_planner_create(2, C.byref(p)) == 0
    for rep in range(2):
        seen = []
        for i in range(6):
            pl = _lib.ScStreamPlan()
            assert lib.sc_planner_push(p, 1, 8192, 0, C.byref(pl)) == 0
            seen.append((pl.n_feat, pl.n_enc_out, pl.n_decode_blocks))
        assert [s[0] for s in seen][:4] == [49, 49, 50, 49]
        assert [s[1] for s in seen][:5] == [0, 0, 0, 24, 16]
        assert [s[2] for s in seen][:5] == [0, 0, 0, 0, 1]
        lib.sc_planner_reset(p, 1)
    lib.sc_planner_destroy(p)


def test_planner_replays_every_golden_call_list():
    """The reference-generated goldens (incl. the live pattern with mid-stream finals) through the host planner:
    feature frames and encoder frames per call must be the reference's."""
    from helpers import GOLDEN_CASES, load_golden
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    for case in GOLDEN_CASES:
        meta, calls, _ = load_golden(case)
        p = C.c_void_p()
        assert lib.sc_planner_create(1, C.byref(p)) == 0
        try:
            for ci, ((s, e, fin), g) in enumerate(zip(meta["calls"], calls)):
                pl = _lib.ScStreamPlan()
                assert lib.sc_planner_push(p, 0, e - s, int(fin), C.byref(pl)) == 0, lib.sc_last_error()
                lite = bool(meta.get("lite"))          # long-utterance goldens hold the shapes as plain numbers
                called = g["called"] if lite else g["feats"] is not None
                assert bool(pl.called) == called, (case, ci)
                if not called:
                    continue
                assert pl.n_feat == (g["n_feat"] if lite else g["feats"].shape[0]), (case, ci)
                assert pl.n_enc_out == (g["n_enc"] if lite else (0 if g["enc"] is None else g["enc"].shape[0])), (case, ci)
        finally:
            lib.sc_planner_destroy(p)


@pytest.mark.parametrize("seed", range(8))
def test_planner_feature_mode_matches_oracle_shapes(seed):
    """Pre-computed feature input (speech2text_streaming.py:438-450): no frontend, process_block on every call."""
    from oracle.encoder import EncoderOracle
    from oracle.speech2text import load_model_dir
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    W, _, _, _ = load_model_dir(model_dir("m_d2"))
    rng = np.random.default_rng(100 + seed)
    n_calls = int(rng.integers(3, 12))
    frames = [int(rng.choice([100, 100, 64, 37, 9, 2, 1, 0, 48])) for _ in range(n_calls)]
    finals = [bool(rng.random() < 0.15) for _ in range(n_calls)]
    finals[-1] = True
    enc = EncoderOracle(W, 2, 4)
    enc_len, processed_block, want = 0, 0, []
    try:
        for n, fin in zip(frames, finals):
            n_enc = 0
            if n >= 3:
                n_enc = enc(torch.zeros(1, n, 80), fin).size(1)
            enc_len += n_enc
            nd, last_T = 0, 0
            while enc_len > 0 and 24 + 16 * processed_block < enc_len:
                last_T = 24 + 16 * processed_block
                nd += 1
                processed_block += 1
            if fin and enc_len > 0:
                nd += 1
                last_T = enc_len
            want.append(dict(called=1, n_feat=n, n_enc_out=n_enc, enc_len=enc_len, n_decode=nd, last_T=last_T))
    except Exception:
        pytest.skip("the reference algorithm itself cannot process this chunk pattern")
    p = C.c_void_p()
    assert lib.sc_planner_create(1, C.byref(p)) == 0
    try:
        for i, (n, fin) in enumerate(zip(frames, finals)):
            pl = _lib.ScStreamPlan()
            assert lib.sc_planner_push_features(p, 0, n, int(fin), C.byref(pl)) == 0, lib.sc_last_error()
            got = dict(called=pl.called, n_feat=pl.n_feat, n_enc_out=pl.n_enc_out, enc_len=pl.enc_len,
                       n_decode=pl.n_decode_blocks, last_T=pl.last_T)
            assert got == want[i], f"call {i} frames={frames} finals={finals}: {got} != {want[i]}"
    finally:
        lib.sc_planner_destroy(p)
