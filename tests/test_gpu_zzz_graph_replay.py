"""GPU: opt-in CUDA-graph replay of the search iteration / encoder stack (engine options graph_decode, graph_encoder)
gives exactly the results of plain launches.

The engine falls back to plain launches when a capture is refused, so every graph run also asserts that graphs were
really replayed (engine counter "graphs_replayed").  scripts/graph_replay_ab.py is the same check as a script with
timings."""
import numpy as np
import pytest
import torch

from helpers import model_dir

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _run(md, graph, dtype, lengths):
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    S = len(lengths)
    st = torch.cuda.Stream()                       # capture is refused on the legacy default stream
    with torch.cuda.stream(st):
        g = StreamGroup(md, n_streams=S, beam_size=5, dtype=dtype, max_seconds=12.0)
        g.set_option("graph_decode", graph & 1)
        g.set_option("graph_encoder", (graph >> 1) & 1)
        audio = [synth_audio(10 + s, n) for s, n in enumerate(lengths)]
        pos, done, beams = [0] * S, [False] * S, []
        while not all(done):
            ids, chunks, fins = [], [], []
            for s in range(S):
                if done[s]:
                    continue
                a = audio[s][pos[s]: pos[s] + 8192]
                fin = pos[s] + 8192 >= lengths[s]
                ids.append(s); chunks.append(a); fins.append(fin)
                pos[s] += 8192
                done[s] = fin
            g.push(ids, chunks, fins)
            beams.append([g.beam(s) for s in ids])
        st.synchronize()
    replayed, failed = g.counter("graphs_replayed"), g.counter("graph_failed")
    g.close()
    if graph:
        assert not failed, "graph capture was refused: the run silently used plain launches"
        assert replayed > 0, "no CUDA graph was replayed"
    else:
        assert replayed == 0
    return beams


@pytest.mark.parametrize("graph", [1, 2, 3])
def test_graph_replay_is_bit_identical_in_fp32(graph):
    md = model_dir("xl_d4")
    lengths = [9 * 16000 + 77, 7 * 16000, 10 * 16000 + 4000]
    for dtype in ("float32_tc", "float32_simt"):
        base = _run(md, 0, dtype, lengths)
        got = _run(md, graph, dtype, lengths)
        for x, y in zip(base, got):
            for a, b in zip(x, y):
                assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3]
                np.testing.assert_array_equal(a[1], b[1])


def test_graph_replay_same_search_in_bf16():
    md = model_dir("xl_d4")
    lengths = [8 * 16000, 6 * 16000 + 500]
    base = _run(md, 0, "bfloat16", lengths)
    got = _run(md, 3, "bfloat16", lengths)
    for x, y in zip(base, got):
        for a, b in zip(x, y):
            assert a[0] == b[0] and a[2] == b[2]


def test_wide_mma_attention_beam20_tracks_cuda_core_attention():
    """BASELINE config 5 (beam 20, bf16): the tiled tensor-core decoder attention (option mma_attention = 2: two m16
    tiles of hypotheses per stream) against the CUDA-core attention the engine uses for beam > 16 by default.
    Decoder log-probs must agree to bf16 accuracy for as long as both searches hold the same beams."""
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("xl_d4")
    S, n = 2, 4 * 16000
    audio = [synth_audio(70 + s, n) for s in range(S)]

    def run(wide):
        g = StreamGroup(md, n_streams=S, beam_size=20, dtype="bfloat16", max_seconds=6.0)
        if wide:
            g.set_option("mma_attention", 2)
        logs, beams = [], []
        for i in range(0, n, 8192):
            fin = i + 8192 >= n
            g.push(list(range(S)), [a[i:i + 8192] for a in audio], [fin] * S)
            logs.append(g.buffer("dlogp").view(-1, 1024)[: S * 20].clone().cpu().numpy())
            beams.append([g.beam(s)[0] for s in range(S)])
        g.close()
        return logs, beams

    a_logs, a_beams = run(False)
    b_logs, b_beams = run(True)
    agree = 0
    for la, lb, ba, bb in zip(a_logs, b_logs, a_beams, b_beams):
        if ba != bb:
            break
        agree += 1
        if np.abs(la).sum() > 0:
            assert np.abs(la - lb).max() < 2e-2
    assert agree >= 4                                   # the first decode blocks take the same decisions
    assert all(len(b) == 20 for b in b_beams[-1]) and [b[0][:4] for b in a_beams[-1]] == [b[0][:4] for b in b_beams[-1]]
