"""GPU: opt-in CUDA-graph replay of the search iteration / encoder stack (engine options graph_decode, graph_encoder)
gives exactly the results of plain launches.

Written after round 1's GPU minutes were spent: xfail(strict=False) until it has run on a device once, and the file
name sorts last so that nothing here can disturb the rest of the suite.  tests/graph_replay_ab.py is the same check as a
script with timings."""
import numpy as np
import pytest
import torch

from helpers import model_dir

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180), pytest.mark.xfail(strict=False, reason="graph replay not yet run on a device")]


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
    g.close()
    return beams


@pytest.mark.parametrize("graph", [1, 2, 3])
def test_graph_replay_is_bit_identical_in_fp32(graph):
    md = model_dir("xl_d4")
    lengths = [9 * 16000 + 77, 7 * 16000, 10 * 16000 + 4000]
    base = _run(md, 0, "float32", lengths)
    got = _run(md, graph, "float32", lengths)
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
