"""GPU: live sessions -- streams that continue after a finalised utterance without reset, and a
pool of sessions stepped with one batched push, against per-session CPU oracles."""
import numpy as np
import pytest

from helpers import model_dir

pytestmark = pytest.mark.gpu


def _chunks(seed, n, size=8192):
    from speechcatcher_b200.synthetic import synth_audio
    a = (synth_audio(seed, n * size) * 32768.0).astype(np.int16)
    return [a[i * size:(i + 1) * size] for i in range(n)]


def test_stream_continues_after_final_without_reset():
    """The live server finalises utterances in the middle of a connection and keeps feeding the same recogniser
    (speechcatcher_server.py:252-270): frontend and encoder restart, the encoder memory and the hypotheses carry on."""
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import StreamGroup
    md = model_dir("m_d2", eos_bias=7.0)
    S, n = 2, 18
    finals = [{5, 9, 17}, {3, 4, 12, 17}]
    grp = StreamGroup(md, n_streams=S, beam_size=5, max_seconds=30.0)
    orc = [OracleSpeech2Text(md, beam_size=5) for _ in range(S)]
    audio = [[c.astype(np.float32) / 32768.0 for c in _chunks(40 + s, n)] for s in range(S)]
    n_final_results = 0
    for i in range(n):
        fins = [i in finals[s] for s in range(S)]
        grp.push(list(range(S)), [audio[s][i] for s in range(S)], fins)
        for s in range(S):
            want = orc[s](audio[s][i], is_final=fins[s], finalize_all=False)
            assert bool(grp.last_plan(s).called) == (orc[s].last_feats is not None)
            if not grp.last_plan(s).called:
                continue
            ys, sc, xp, pidx = grp.beam(s)
            assert ys == [list(h.yseq) for h in orc[s].hyps], f"stream {s} call {i}"
            assert xp == [list(h.xpos) for h in orc[s].hyps], f"stream {s} call {i}"
            np.testing.assert_allclose(sc, [h.score for h in orc[s].hyps], atol=2e-3, rtol=0)
            got = grp.results(s, fins[s], False)
            assert [r[2] for r in got] == [r[2] for r in want]
            n_final_results += int(fins[s] and len(got) > 0)
    assert n_final_results >= 3


class _OracleFacade:
    """`speech2text(speech=, is_final=)` over the CPU oracle (what a reference session would call)."""

    def __init__(self, md):
        from oracle.speech2text import OracleSpeech2Text
        self.o = OracleSpeech2Text(md, beam_size=3)

    def reset(self):
        self.o.reset()

    def __call__(self, speech, is_final=False, finalize_all=False):
        return self.o(np.asarray(speech, np.float32), is_final=is_final, finalize_all=finalize_all)


@pytest.mark.parametrize("reset_on_finalize", [False, True])
def test_pool_sessions_with_one_batched_push_match_oracle_sessions(reset_on_finalize):
    from speechcatcher_b200.live import LiveSession, StreamPool, process_many
    md = model_dir("m_d2", eos_bias=7.0)
    pool = StreamPool(md, beam_size=3, pool_size=3, max_seconds=30.0)
    views = [pool.acquire() for _ in range(3)]
    assert all(v is not None for v in views) and pool.acquire() is None
    assert len({id(v.group) for v in views}) == 1 and sorted(v.stream_id for v in views) == [0, 1, 2]
    kw = dict(finalize_update_iters=3, vosk_output_format=True, reset_on_finalize=reset_on_finalize)
    sessions = [LiveSession(v, **kw) for v in views]
    ref_sessions = [LiveSession(_OracleFacade(md), **kw) for _ in views]
    scripts = [_chunks(60 + k, 16) for k in range(3)]
    scripts[1][7] = '{"eof" : 1}'
    scripts[2][4] = np.zeros(0, np.int16)
    n_results = 0
    for t in range(16):
        got = process_many(sessions, [scripts[k][t] for k in range(3)])
        want = [ref_sessions[k].process_audio_chunk(scripts[k][t]) for k in range(3)]
        assert got == want, t
        n_results += sum(1 for g in got if isinstance(g, dict) and g.get("result"))
    assert n_results >= 3
    # a single session stepped on its own goes through the facade call and gives the same outputs
    pool.release(views[0])
    v = pool.acquire()
    solo, ref = LiveSession(v, **kw), LiveSession(_OracleFacade(md), **kw)
    for t in range(8):
        assert solo.process_audio_chunk(scripts[0][t]) == ref.process_audio_chunk(scripts[0][t])
