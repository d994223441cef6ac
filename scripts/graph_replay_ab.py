"""Round-2 check for the opt-in CUDA-graph replay (engine options "graph_decode" / "graph_encoder"): NOT collected by
pytest (the options were written without GPU time left in round 1 and have never run on a device).

    python scripts/graph_replay_ab.py            # on a B200 box

1. parity: three ragged streams on an explicit (non-default) CUDA stream with graphs on must give exactly the beams of
   the same engine with graphs off (fp32 mode is bit-exact run to run) and of the CPU oracle;
2. prints whether the graphs were actually captured (launch counters) and the time of both variants.
Then: `SCB_BENCH_GRAPH=3 scripts/bench_value.sh` against `scripts/bench_value.sh` for the throughput effect; if it
holds, make it the default in bench.py and add this file's parity part to tests/test_gpu_multistream.py.
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def run(md, graph, dtype, lengths, seconds_cap=12.0):
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.synthetic import synth_audio
    S = len(lengths)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g = StreamGroup(md, n_streams=S, beam_size=5, dtype=dtype, max_seconds=seconds_cap)
        g.set_option("graph_decode", graph & 1)
        g.set_option("graph_encoder", (graph >> 1) & 1)
        audio = [synth_audio(10 + s, n) for s, n in enumerate(lengths)]
        pos, done, beams = [0] * S, [False] * S, []
        t0 = time.perf_counter()
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
        dt = time.perf_counter() - t0
    return beams, dt, g.total_launches


def main():
    from helpers import model_dir
    md = model_dir("xl_d4")
    lengths = [9 * 16000 + 77, 7 * 16000, 10 * 16000 + 4000]
    for dtype in ("float32", "bfloat16"):
        base, t_base, l_base = run(md, 0, dtype, lengths)
        for graph in (1, 2, 3):
            got, t, l = run(md, graph, dtype, lengths)
            same = all(a[0] == b[0] and a[2] == b[2] and np.allclose(a[1], b[1], atol=0 if dtype == "float32" else 1e-3)
                       for x, y in zip(base, got) for a, b in zip(x, y))
            print(f"{dtype} graph={graph}: identical beams={same}  time {t:.3f}s vs {t_base:.3f}s  "
                  f"launch counter {l} vs {l_base}")
            assert same, "graph replay changed the results"
    print("graph replay parity ok")


if __name__ == "__main__":
    main()
