"""GPU: several ragged streams batched in one StreamGroup against the live CPU oracle (one oracle
object per stream), fp32 mode: n-best tokens, timestamps and order exact; scores 2e-3."""
import numpy as np
import pytest
import torch

from helpers import model_dir

pytestmark = pytest.mark.gpu


def _run(arch, beam, lengths, chunk_of, use_bbd=False, kind="noise", dtype="float32"):
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir(arch)
    S = len(lengths)
    audio = [synth_audio(10 + s, n, kind) for s, n in enumerate(lengths)]
    grp = StreamGroup(md, n_streams=S, beam_size=beam, device="cuda:0", dtype=dtype, use_bbd=use_bbd,
                      max_chunk=max(chunk_of.values()), max_seconds=max(lengths) / 16000 + 2)
    orc = [OracleSpeech2Text(md, beam_size=beam, use_bbd=use_bbd) for _ in range(S)]
    pos = [0] * S
    done = [False] * S
    step = 0
    while not all(done):
        ids, chunks, fins = [], [], []
        for s in range(S):
            if done[s] or (step % 3 == 2 and s % 2 == 1):      # some streams skip some pushes
                continue
            c = chunk_of[s]
            a = audio[s][pos[s]: pos[s] + c]
            fin = pos[s] + c >= lengths[s]
            ids.append(s); chunks.append(a); fins.append(fin)
            pos[s] += c
            done[s] = fin
        step += 1
        if not ids:
            continue
        grp.push(ids, chunks, fins)
        for s, a, fin in zip(ids, chunks, fins):
            want = orc[s](a, is_final=fin, finalize_all=fin)
            plan = grp.last_plan(s)
            assert bool(plan.called) == (orc[s].last_feats is not None)
            if not plan.called:
                continue
            ys, sc, xp, pidx = grp.beam(s)
            hyps = orc[s].hyps
            assert ys == [list(h.yseq) for h in hyps], f"stream {s} step {step}"
            assert xp == [list(h.xpos) for h in hyps], f"stream {s} step {step}"
            np.testing.assert_allclose(sc, [h.score for h in hyps], atol=2e-3, rtol=0)
            assert pidx == orc[s].search.process_idx
            got = grp.results(s, fin, fin)
            assert [r[2] for r in got] == [r[2] for r in want]
    return grp


def test_three_ragged_streams_m_d2():
    _run("m_d2", 5, [5 * 16000 + 77, 3 * 16000, 6 * 16000 + 4000], {0: 8192, 1: 8192, 2: 8192})


def test_mixed_chunk_sizes_xl_d4_beam10():
    _run("xl_d4", 10, [4 * 16000, 4 * 16000 + 123, 20000, 3 * 16000], {0: 8192, 1: 4096, 2: 8192, 3: 6000}, kind="tones")


def test_bbd_streams():
    _run("m_d2", 5, [4 * 16000, 4 * 16000 + 999], {0: 8192, 1: 8192}, use_bbd=True)


def test_reset_reuses_stream_slot():
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import Speech2TextStreaming
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2")
    gpu = Speech2TextStreaming(md, beam_size=5, device="cuda:0")
    for utt in range(2):
        audio = synth_audio(40 + utt, 3 * 16000 + 500)
        orc = OracleSpeech2Text(md, beam_size=5)
        gpu.reset()
        for i in range(0, len(audio), 8192):
            fin = i + 8192 >= len(audio)
            a = gpu(audio[i:i + 8192], is_final=fin, finalize_all=fin)
            b = orc(audio[i:i + 8192], is_final=fin, finalize_all=fin)
            assert [r[2] for r in a] == [r[2] for r in b]
            assert [r[0] for r in a] == [r[0] for r in b]
        assert gpu.beam_state[0] == [list(h.yseq) for h in orc.hyps]


def test_deferred_decoding_same_final_results():
    """Deferred (lazy) decoding only reorders when queued decode blocks run: every stream's final n-best must
    equal the oracle's exactly (fp32 mode), even though non-final beams lag behind."""
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2")
    S, n = 6, 5 * 16000 + 321
    audio = [synth_audio(70 + s, n) for s in range(S)]
    grp = StreamGroup(md, n_streams=S, beam_size=5, device="cuda:0", max_seconds=8)
    grp.set_option("lazy_threshold", 4)
    total_steps = 0
    for i in range(0, n, 8192):
        fin = i + 8192 >= n
        st = grp.push(list(range(S)), [a[i:i + 8192] for a in audio], [fin] * S)
        total_steps += st.n_decode_steps
    for s in range(S):
        orc = OracleSpeech2Text(md, beam_size=5)
        for i in range(0, n, 8192):
            fin = i + 8192 >= n
            want = orc(audio[s][i:i + 8192], is_final=fin, finalize_all=fin)
        ys, sc, xp, pidx = grp.beam(s)
        assert ys == [list(h.yseq) for h in orc.hyps]
        assert xp == [list(h.xpos) for h in orc.hyps]
        np.testing.assert_allclose(sc, [h.score for h in orc.hyps], atol=2e-3, rtol=0)
        assert [r[2] for r in grp.results(s, True, True)] == [r[2] for r in want]
    bulk = grp.results_all(True, True)                 # one bulk read-back == per-stream reads
    for s in range(S):
        one = grp.results(s, True, True)
        assert [r[:4] for r in bulk[s]] == [r[:4] for r in one]
        for (*_, hb), (*_, ho) in zip(bulk[s], one):
            assert hb["score"] == ho["score"]
            assert hb["yseq"].tolist() == ho["yseq"].tolist() and hb["xpos"].tolist() == ho["xpos"].tolist()
    assert total_steps > 0


def test_beam20_fp32_and_bf16_smoke():
    """BASELINE config 5 shape (beam 20): exact parity in fp32; the bf16 mode (CUDA-core attention for beam > 16)
    must run and produce a full beam of finite scores."""
    grp = _run("xl_d4", 20, [3 * 16000 + 50, 3 * 16000], {0: 8192, 1: 8192})
    assert grp.beam_size == 20
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    g16 = StreamGroup(model_dir("xl_d4"), n_streams=2, beam_size=20, device="cuda:0", dtype="bfloat16", max_seconds=6)
    n = 3 * 16000
    audio = [synth_audio(90 + s, n) for s in range(2)]
    for i in range(0, n, 8192):
        fin = i + 8192 >= n
        g16.push([0, 1], [a[i:i + 8192] for a in audio], [fin, fin])
    for s in range(2):
        ys, sc, xp, _ = g16.beam(s)
        assert len(ys) == 20 and all(np.isfinite(sc)) and len(ys[0]) > 3


def test_sharded_group_matches_oracle():
    """Two engines on one GPU, each on its own CUDA stream and host thread, deferred decoding + encoder overlap:
    every stream's final n-best equals the oracle's (fp32 mode)."""
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200.sharded_group import ShardedStreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2")
    S, n = 4, 4 * 16000 + 777
    audio = np.stack([synth_audio(80 + s, n) for s in range(S)])
    sg = ShardedStreamGroup(md, S, 2, device="cuda:0", beam_size=5, max_seconds=7)
    sg.set_option("lazy_threshold", 2)
    ids = np.arange(2, dtype=np.int32)

    def work(i, g, lo, hi):
        g.reset()
        for c in range(0, n, 8192):
            fin = c + 8192 >= n
            g.push(list(ids), [audio[s, c:c + 8192] for s in range(lo, hi)], [fin] * (hi - lo))

    sg.run_pass(work)
    sg.synchronize()
    for s in range(S):
        orc = OracleSpeech2Text(md, beam_size=5)
        for c in range(0, n, 8192):
            fin = c + 8192 >= n
            want = orc(audio[s, c:c + 8192], is_final=fin, finalize_all=fin)
        ys, sc, xp, _ = sg.beam(s)
        assert ys == [list(h.yseq) for h in orc.hyps], f"stream {s}"
        assert xp == [list(h.xpos) for h in orc.hyps]
        assert [r[2] for r in sg.results(s, True, True)] == [r[2] for r in want]
    sg.close()


def test_odd_chunk_sizes_against_oracle():
    """Chunk sizes that exercise the buffering corners: tiny chunks (frontend emits 5-8 frames, the encoder buffers
    for many calls), chunks that are not multiples of the hop, and chunks larger than the default."""
    _run("m_d2", 5, [3 * 16000, 3 * 16000 + 1234, 4 * 16000], {0: 1600, 1: 3000, 2: 12000})


def test_batch_invariance_large_group():
    """Size-independent property at a scale the CPU oracle cannot cover: a stream decoded inside a 24-stream group
    (deferred decoding, two shards, ragged lengths) gives exactly the beam it gives when decoded alone."""
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.sharded_group import ShardedStreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("xl_d4")
    S = 24
    lens = [int((8 + (s % 5)) * 16000 + 137 * s) for s in range(S)]
    audio = [synth_audio(200 + s, n) for s, n in enumerate(lens)]
    sg = ShardedStreamGroup(md, S, 2, device="cuda:0", beam_size=10, max_seconds=15)
    sg.set_option("lazy_threshold", 10)

    def work(i, g, lo, hi):
        g.reset()
        pos = 0
        live = list(range(lo, hi))
        while live:
            ids, chunks, fins = [], [], []
            for s in list(live):
                a = audio[s][pos: pos + 8192]
                fin = pos + 8192 >= lens[s]
                ids.append(s - lo); chunks.append(a); fins.append(fin)
                if fin:
                    live.remove(s)
            g.push(ids, chunks, fins)
            pos += 8192

    sg.run_pass(work)
    sg.synchronize()
    solo = StreamGroup(md, n_streams=1, beam_size=10, device="cuda:0", max_seconds=15)
    for s in (0, 5, 11, 12, 17, 23):
        solo.reset()
        for pos in range(0, lens[s], 8192):
            solo.push([0], [audio[s][pos: pos + 8192]], [pos + 8192 >= lens[s]])
        a, b = sg.beam(s), solo.beam(0)
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3], f"stream {s}"
        np.testing.assert_array_equal(np.asarray(a[1]), np.asarray(b[1]))     # fp64 scores bit-identical
    sg.close()


def test_failed_push_leaves_every_stream_untouched():
    """A push that is rejected after some of its streams were already planned (here: stream 1 sends a final chunk the
    reference itself cannot process) must roll the host planner back: stream 0 then continues exactly as if the failed
    call had never happened.  Also: duplicate stream ids in one push are rejected."""
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2")
    n = 4 * 16000 + 321
    audio = synth_audio(77, n)
    grp = StreamGroup(md, n_streams=2, beam_size=5, device="cuda:0", max_seconds=8.0)
    orc = OracleSpeech2Text(md, beam_size=5)
    pos = 0
    k = 0
    while pos < n:
        a = audio[pos: pos + 8192]
        fin = pos + 8192 >= n
        if k in (1, 4):
            with pytest.raises(RuntimeError):          # 100 samples, final, nothing buffered: conv2d would raise upstream
                grp.push([0, 1], [a, np.zeros(100, np.float32)], [fin, True])
            with pytest.raises(RuntimeError):
                grp.push([0, 0], [a, a], [fin, fin])
        grp.push([0], [a], [fin])
        orc(a, is_final=fin, finalize_all=fin)
        if grp.last_plan(0).called:
            ys, sc, xp, pidx = grp.beam(0)
            assert ys == [list(h.yseq) for h in orc.hyps], f"chunk {k}"
            assert xp == [list(h.xpos) for h in orc.hyps]
            assert pidx == orc.search.process_idx
        pos += 8192
        k += 1
    # stream 1 was never advanced by the failed pushes: a regular utterance on it still matches
    orc1 = OracleSpeech2Text(md, beam_size=5)
    b = synth_audio(78, 2 * 16000 + 999)
    for i in range(0, len(b), 8192):
        fin = i + 8192 >= len(b)
        grp.push([1], [b[i:i + 8192]], [fin])
        orc1(b[i:i + 8192], is_final=fin, finalize_all=fin)
    assert grp.beam(1)[0] == [list(h.yseq) for h in orc1.hyps]
