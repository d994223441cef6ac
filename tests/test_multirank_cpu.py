"""CPU, world_size 2 over gloo: the N>1 host logic (stream sharding + the end-of-run result gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speechcatcher_b200.sharding import owner_of, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speechcatcher_b200.sharding import gather_results
    lo, hi = shard_range(n_streams, world, rank)
    local = {s: dict(tokens=[s, s + 1], rank=rank) for s in range(lo, hi)}
    allr = gather_results(local)
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # same pattern as bench.py's max-over-ranks time
    q.put((rank, sorted(allr), [allr[s]["rank"] for s in sorted(allr)], t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [7, 512])
def test_shard_and_gather_world2(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ids, owners, mx in got:
        assert ids == list(range(n_streams))
        assert owners == [owner_of(s, n_streams, 2) for s in range(n_streams)]
        assert mx == float(shard_range(n_streams, 2, 0)[1])


def test_shard_range_partitions():
    for n in (1, 5, 256, 2048):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = shard_range(n, w, r)
                cover += list(range(lo, hi))
            assert cover == list(range(n))


def _recognize_worker(rank, world, port, case_name, q):
    """Long-file recognition sharded by segment (config 4 on N GPUs): scripted recogniser, gloo gather."""
    import json
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import GOLDEN
    from oracle.gen_golden_recognize import case_audio
    from oracle.scripted_backend import ScriptedGroup
    from speechcatcher_b200.recognize import recognize
    case = next(c for c in json.loads((GOLDEN / "recognize.json").read_text())["recognize"] if c["name"] == case_name)
    group = ScriptedGroup(3)
    text, aux = recognize(group, case_audio(case["seed"], case["n"]), 16000, chunk_length=case["chunk"],
                          segments=[tuple(x) for x in case["segments"]], shard=(rank, world))
    n_calls = sum(len(p) for p in group.pushes)
    q.put((rank, text == case["text"], aux == case["aux"], n_calls, len(case["calls"])))
    dist.destroy_process_group()


def test_long_file_sharded_by_segment_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_recognize_worker, args=(r, 2, port, "many_segments_chunk4000", q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok_text and ok_aux for _, ok_text, ok_aux, _, _ in got)        # every rank holds the full transcript
    assert sum(n for _, _, _, n, _ in got) == got[0][4]                        # the calls were split, not duplicated
    assert all(0 < n < got[0][4] for _, _, _, n, _ in got)


def _beams_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from speechcatcher_b200.sharding import gather_beams
    S, B, L = 5, 3, 16
    rng = np.random.default_rng(rank)
    ctl = np.stack([np.full(S, B), rng.integers(1, L, S)], axis=1).astype(np.int32)
    ys = rng.integers(0, 1024, (S, B, L)).astype(np.int32)
    xp = rng.integers(0, 1500, (S, B, L)).astype(np.int32)
    sc = rng.standard_normal((S, B))
    g_ctl, g_ys, g_xp, g_sc = gather_beams(ctl, ys, xp, sc)
    ok = g_ctl.shape == (world * S, 2) and g_ys.shape == (world * S, B, L) and g_sc.dtype == np.float64
    ok = ok and np.array_equal(g_ys[rank * S:(rank + 1) * S], ys) and np.array_equal(g_sc[rank * S:(rank + 1) * S], sc)
    other = 1 - rank
    ro = np.random.default_rng(other)
    ro.integers(1, L, S)
    ok = ok and np.array_equal(g_ys[other * S:(other + 1) * S], ro.integers(0, 1024, (S, B, L)).astype(np.int32))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_fixed_width_beam_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_beams_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in got)
