"""BASELINE config 4 on one GPU: a long file -> device segmentation (N2) -> segments as concurrent streams (N1).

    python scripts/bench_longfile.py [--seconds 3600] [--arch l] [--streams 64] [--dtype bfloat16] [--beam 10]

Prints one JSON line: segmentation time (H2D of the int16 file + fp64 energy/smoothing kernels + D2H + host cut search,
CUDA-event timed where it runs on the device), the numpy oracle's time for the same curve on a bounded sample, and the
wall time / audio-s/s of the whole file through `speechcatcher_b200.recognize.recognize`.
Random-init weights and synthetic audio (noise bursts separated by pauses); no network, no checkpoints.
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--arch", default="l")
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--beam", type=int, default=10)
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--cpu-sample-seconds", type=float, default=300.0)
    ap.add_argument("--skip-decode", action="store_true")
    a = ap.parse_args()

    import torch
    from oracle.gen_golden_endpointing import pause_audio
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.recognize import plan_segments, recognize
    from speechcatcher_b200.simple_endpointing import BeamSearch, cap_segments, smoothed_energy
    from speechcatcher_b200.synthetic import make_model_dir

    audio = pause_audio(4, a.seconds)
    out = {"workload": f"{a.seconds:.0f} s synthetic file, {a.arch} arch, beam {a.beam}, {a.dtype}, {a.streams} streams"}

    smoothed_energy(audio[: 16000 * 5])                      # warm-up (module load, constant tables)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0 = time.perf_counter()
    ev[0].record()
    curve = smoothed_energy(audio)
    ev[1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    search = BeamSearch(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0, energy_weight=1.0)
    segs = cap_segments(search.search(curve, len(curve)), 180)
    t2 = time.perf_counter()
    out["segmenter"] = {"frames": len(curve), "segments": len(segs),
                        "energy_ms_device_incl_copies": ev[0].elapsed_time(ev[1]), "energy_ms_wall": (t1 - t0) * 1e3,
                        "cut_search_ms_host": (t2 - t1) * 1e3,
                        "longest_segment_s": max(e - s for s, e in segs) / 100.0}

    from oracle.endpointing import CutSearchOracle, smoothed_energy as cpu_energy
    n_cpu = int(min(a.cpu_sample_seconds, a.seconds) * 16000)
    t0 = time.perf_counter()
    cpu_curve = cpu_energy(audio[:n_cpu])
    t1 = time.perf_counter()
    CutSearchOracle(beam_size=10, ideal_segment_len=6000, step=10, len_reward_weight=12.0,
                    energy_weight=1.0).search(cpu_curve, len(cpu_curve))
    t2 = time.perf_counter()
    out["segmenter_cpu_oracle"] = {"sample_seconds": n_cpu / 16000, "energy_ms": (t1 - t0) * 1e3,
                                   "cut_search_ms": (t2 - t1) * 1e3, "cores": 1}

    if not a.skip_decode:
        with tempfile.TemporaryDirectory() as td:
            md = make_model_dir(td, a.arch, seed=0)
            plan = plan_segments(len(audio), 16000, segs, 8192)
            longest = max(e - s for s, e in plan.seconds)
            group = StreamGroup(md, n_streams=min(a.streams, plan.n_segments), beam_size=a.beam, dtype=a.dtype,
                                max_seconds=longest + 3.0)
            recognize(group, audio[: 16000 * 20], 16000, segments=[])          # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            text, aux = recognize(group, audio, 16000, segments=segs)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out["decode"] = {"segments": plan.n_segments, "streams": group.n_streams, "wall_s": dt,
                             "audio_s_per_s": a.seconds / dt, "paragraphs": len(aux),
                             "tokens": sum(len(p["tokens"]) for p in aux), "launches": group.total_launches}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
