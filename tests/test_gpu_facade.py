"""GPU: the drop-in facade keeps the reference's call protocol (speechcatcher/speech2text_streaming.py:29-621)."""
import numpy as np
import pytest
import torch

from helpers import model_dir

pytestmark = pytest.mark.gpu


def test_facade_surface_and_tuple_shape():
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import Speech2TextStreaming, create_streaming_interface
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2")
    s2t = create_streaming_interface(md, beam_size=5, device="cuda")
    assert isinstance(s2t, Speech2TextStreaming)
    for attr in ("model", "beam_search", "beam_size", "beam_state", "frontend_states", "processed_frames", "mean", "std", "token_list",
                 "win_length", "hop_length", "reset", "recognize", "recognize_stream", "n_best_hypotheses",
                 "get_best_hypothesis"):
        assert hasattr(s2t, attr), attr
    assert s2t.n_best_hypotheses == 5 and s2t.win_length == 400 and s2t.hop_length == 160
    assert s2t.model is not None and s2t.beam_search.use_bbd is False and s2t.beam_search.weights["ctc"] == 0.3
    audio = synth_audio(7, 3 * 16000 + 99)
    chunks = [audio[i:i + 8192] for i in range(0, len(audio), 8192)]
    # recognize_stream: last chunk is_final, finalize_all False -> only hypotheses ending in <eos> (id 1023) are returned
    orc = OracleSpeech2Text(md, beam_size=5)
    want = None
    for i, c in enumerate(chunks):
        want = orc(c, is_final=(i == len(chunks) - 1))
    got = s2t.recognize_stream(chunks)
    assert [r[2] for r in got] == [r[2] for r in want]
    # torch tensors and the CLI's extra keyword are accepted; results are ESPnet-shaped 5-tuples
    s2t.reset()
    res = None
    for i, c in enumerate(chunks):
        fin = i == len(chunks) - 1
        res = s2t(torch.from_numpy(c), is_final=fin, finalize_all=fin, always_assemble_hyps=True)
    assert len(res) == 5
    text, tokens, ids, pos, hyp = res[0]
    assert isinstance(text, str) and tokens == [str(t) for t in ids] and len(pos) == len(ids)
    assert text == " ".join(tokens)
    assert set(hyp) == {"yseq", "score", "xpos"} and hyp["yseq"][0] == 1023
    # token positions are encoder frame indices of the block that emitted the token: non-decreasing
    assert all(b >= a for a, b in zip(pos, pos[1:]))
    best = s2t.get_best_hypothesis()
    assert best[2] == ids
    # recognize(): one-shot, is_final without finalize_all
    one = s2t.recognize(audio[:20000])
    orc2 = OracleSpeech2Text(md, beam_size=5)
    assert [r[2] for r in one] == [r[2] for r in orc2(audio[:20000], is_final=True)]


def test_errors_are_loud():
    from speechcatcher_b200 import Speech2TextStreaming, StreamGroup
    md = model_dir("m_d2")
    with pytest.raises(RuntimeError):
        StreamGroup(md, n_streams=1, device="cpu")
    s2t = Speech2TextStreaming(md, beam_size=5, device="cuda:0", max_chunk=8192)
    s2t(np.zeros(4000, np.float32))
    with pytest.raises((RuntimeError, ValueError)):
        s2t(np.zeros(20000, np.float32))              # larger than max_chunk in the middle of an utterance
    with pytest.raises(ValueError):
        s2t(np.zeros((1, 1, 10, 80), np.float32))     # 1-D waveforms, 2-D features or 3-D batched features only
    with pytest.raises(RuntimeError):
        StreamGroup(md, n_streams=1, beam_size=64)    # beam out of range
