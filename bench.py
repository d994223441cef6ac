#!/usr/bin/env python
"""Benchmark of the multi-stream streaming decode path (BASELINE.json metric: audio-sec/sec, RTFx).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle port of the reference path

Workload (BASELINE.json configs[1]): de_streaming_transformer_xl architecture with random-init
weights, 256 concurrent streams per GPU, beam 10, 60 s of synthetic 16 kHz audio per stream fed in
8192-sample chunks, last chunk final.  One "step" = decoding all streams' 60 s (118 pushes).
Streams are independent, so N GPUs run N x 256 streams with no data-path collective (weak scaling);
torch.distributed (NCCL) is used only for the barrier, the max-over-ranks timing and result checks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

CHUNK = 8192
SR = 16000


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baseline (the reference on host cores)
# oracle/_ref (scripts/make_oracle_ref.sh, git-ignored, travels with gpurun) is the UNMODIFIED reference package: when
# it is present the CPU legs time the reference's own Speech2TextStreaming (kind "reference"), otherwise the oracle
# port of the same path (kind "port").  Both legs of a run (`--impl reference` and `cpu_baseline`) use the same
# sample length, a function of the number of passes only.
_ORACLE_CACHE = {}
REF_DIR = REPO / "oracle" / "_ref"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_kind():
    return "reference" if (REF_DIR / "speechcatcher" / "speech2text_streaming.py").exists() else "port"


def _cpu_worker(args):
    md, stream, n_samples, beam, kind = args
    import contextlib
    import io
    import logging
    import torch
    torch.set_num_threads(1)
    logging.disable(logging.CRITICAL)
    from speechcatcher_b200.synthetic import synth_audio
    audio = synth_audio(stream, n_samples)
    key = (md, beam, kind)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):     # the reference prints debug lines
        if key not in _ORACLE_CACHE:                  # one model load per worker process, reused across steps
            if kind == "reference":
                if str(REF_DIR) not in sys.path:
                    sys.path.insert(0, str(REF_DIR))
                from speechcatcher.speech2text_streaming import Speech2TextStreaming as RefS2T
                _ORACLE_CACHE[key] = RefS2T(md, beam_size=beam, ctc_weight=0.3, device="cpu", dtype="float32", use_bbd=False)
            else:
                from oracle.speech2text import OracleSpeech2Text
                _ORACLE_CACHE[key] = OracleSpeech2Text(md, beam_size=beam, ctc_weight=0.3)
        o = _ORACLE_CACHE[key]
        o.reset()
        t0 = time.perf_counter()
        lat = []
        for i in range(0, n_samples, CHUNK):
            fin = i + CHUNK >= n_samples
            t1 = time.perf_counter()
            o(audio[i:i + CHUNK], is_final=fin, finalize_all=fin)
            lat.append(time.perf_counter() - t1)
        dt = time.perf_counter() - t0
    return dt, lat


class CpuBaseline:
    """`procs` single-threaded worker processes, one stream each per pass (mirrors the reference CLI's process pool,
    speechcatcher.py:481-497, 816-819).  A pass returns (aggregate audio-s/s, p50 per-chunk latency in ms)."""

    def __init__(self, md, beam, procs):
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        self.md, self.beam, self.procs, self.kind = str(md), beam, procs, cpu_kind()
        self.pool = ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("spawn"))

    def run(self, sample_seconds, first_stream=0):
        n = int(sample_seconds * SR)
        jobs = [(self.md, first_stream + i, n, self.beam, self.kind) for i in range(self.procs)]
        res = list(self.pool.map(_cpu_worker, jobs))
        lat = [x for _, l in res for x in l]
        compute = max(r[0] for r in res)          # excludes process start-up / model load
        return self.procs * sample_seconds / compute, 1000.0 * statistics.median(lat)

    def describe(self, sample_seconds):
        what = ("the reference's own Speech2TextStreaming (oracle/_ref, unmodified)" if self.kind == "reference"
                else "oracle port of the reference path (oracle/_ref absent)")
        return (f"{self.procs} streams x {sample_seconds:g} s (first seconds of the workload's streams), one single-threaded "
                f"process per stream, {what}; the reference's per-step cost grows with utterance length, so a sample "
                f"shorter than the workload's 60 s over-states its throughput")

    def close(self):
        self.pool.shutdown(wait=True)


def cpu_sample_seconds(n_passes, budget_s=240.0):
    """Audio seconds per stream of one CPU pass, the same in both CPU legs of a run: 20 s when the
    passes fit the time budget, shorter otherwise (one single-threaded process decodes roughly 0.5 audio-seconds per
    second on this workload, slower as the utterance grows)."""
    per_pass = budget_s / max(1, n_passes)
    return float(min(20.0, max(4.0, 0.45 * (per_pass - 2.0))))


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="xl")
    ap.add_argument("--streams", type=int, default=256, help="streams per GPU")
    ap.add_argument("--beam", type=int, default=10)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--dtype", default=os.environ.get("SCB_BENCH_DTYPE", "float32_tc"),
                    choices=["float32_tc", "float32_simt", "float32", "bfloat16"],
                    help="float32_tc (default): fp32 results with every Linear on the tcgen05 tensor cores as a split-fp16 "
                         "GEMM (parity mode, n-best identical to the reference on the goldens); float32_simt: the same on "
                         "the CUDA cores; bfloat16: bf16 operands (fast, but random-init weights flip the n-best)")
    ap.add_argument("--no-fp32", action="store_true",
                    help="skip the float32_simt pass (its throughput, and the `parity` agreement of the benched mode with it)")
    ap.add_argument("--no-extra-rooflines", action="store_true",
                    help="skip the per-kernel roofline passes on a dedicated single 256-stream group")
    ap.add_argument("--cpu-sample-seconds", type=float, default=None,
                    help="audio seconds per stream of the CPU sample (default: sized so the CPU leg takes ~2.5 min)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-kernel", default="auto")
    ap.add_argument("--breakdown", action="store_true", help="always run the all-kernel timing pass")
    ap.add_argument("--shards", type=int, default=int(os.environ.get("SCB_BENCH_SHARDS", "2")),
                    help="split the GPU's streams into this many concurrently driven groups (own CUDA stream + host thread)")
    ap.add_argument("--graph", type=int, default=int(os.environ.get("SCB_BENCH_GRAPH", "0")),
                    help="CUDA-graph replay (experimental, off by default): 1 = search iteration, 2 = encoder stack, 3 = both")
    ap.add_argument("--enc-sms", type=int, default=int(os.environ.get("SCB_BENCH_ENC_SMS", "0")),
                    help="SM partition of the encoder stream (green context): the frontend + encoder kernels of a push run on "
                         "~N SMs, the rest stays free for the search chain; 0 = no partition")
    ap.add_argument("--lazy", type=int, default=int(os.environ.get("SCB_BENCH_LAZY", "-1")),
                    help="deferred-decode threshold (streams); 0 = strict per-push decoding; -1 = streams - streams/32")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = (f"de_streaming_transformer_{args.arch} arch (random-init), {args.streams} streams/GPU, beam {args.beam}, "
                f"{args.seconds:g} s synthetic 16 kHz audio/stream, 8192-sample chunks, ctc_weight 0.3, use_bbd False")

    from speechcatcher_b200.synthetic import make_model_dir, synth_audio
    tmp = tempfile.TemporaryDirectory(prefix=f"scb200_bench_r{rank}_")
    md = make_model_dir(Path(tmp.name) / args.arch, args.arch, seed=0)

    # ------------------------------------------------------------------ reference arm: CPU oracle port
    if args.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, str(REPO))
        procs = cores
        sample_s = args.cpu_sample_seconds or cpu_sample_seconds(args.warmup + args.steps)
        cpu = CpuBaseline(md, args.beam, procs)
        vals, p50s = [], []
        for i in range(args.warmup + args.steps):
            v, p50 = cpu.run(sample_s)
            if i >= args.warmup:
                vals.append(v); p50s.append(p50)
        cpu.close()
        v = float(np.mean(vals))
        sample = cpu.describe(sample_s)
        line = {"impl": "reference", "metric": "audio-sec/sec (RTFx)", "value": v, "unit": "audio-s/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * procs * sample_s / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "inputs": "host"},
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": procs, "kind": cpu.kind, "sample": sample,
                                 "p50_chunk_ms": float(np.mean(p50s)), "cpu_model": cpu_model()},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ this repo's CUDA path
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from speechcatcher_b200 import StreamGroup

    S = args.streams
    n_samples = int(args.seconds * SR)
    n_chunks = (n_samples + CHUNK - 1) // CHUNK
    host = torch.empty(S, n_chunks * CHUNK, dtype=torch.float32).pin_memory()
    host.zero_()
    for s in range(S):
        host[s, :n_samples] = torch.from_numpy(synth_audio(rank * S + s, n_samples))
    from speechcatcher_b200.sharded_group import ShardedStreamGroup
    G = max(1, args.shards)
    assert S % G == 0, "--streams must be a multiple of --shards"
    per = S // G
    lazy = args.lazy if args.lazy >= 0 else max(1, per - per // 32)

    def make_groups(dtype):
        kw = dict(beam_size=args.beam, ctc_weight=0.3, dtype=dtype, use_bbd=False, max_chunk=CHUNK,
                  max_seconds=args.seconds + 1.0)
        if G == 1:
            g = StreamGroup(md, n_streams=S, device=dev, own_stream=True, **kw)     # own high-priority stream
            g.set_option("lazy_threshold", lazy)
            sg_, shards = None, [g]
        else:
            sg_ = ShardedStreamGroup(md, S, G, device=dev, **kw)
            sg_.set_option("lazy_threshold", lazy)
            shards = sg_.shards
        if args.graph:
            for g in shards:
                g.set_option("graph_decode", args.graph & 1)
                g.set_option("graph_encoder", (args.graph >> 1) & 1)
        if args.enc_sms:
            for g in shards:
                g.set_option("encoder_sms", args.enc_sms)
        return sg_, shards

    sg, groups = make_groups(args.dtype)
    groups_enc_sms = groups[0].counter("encoder_sms")      # SMs the driver really provisioned for the encoder stream (0 = none)
    grp = groups[0]                                   # kernel timing / roofline are taken on shard 0
    resident = host.to(dev)                           # inputs resident in HBM for `value`
    # e2e inputs: what a caller hands over per push -- one pinned [streams, chunk] buffer per chunk
    host_chunks = [host[:, c * CHUNK:(c + 1) * CHUNK].contiguous().pin_memory() for c in range(n_chunks)]
    ids = np.arange(per, dtype=np.int32)
    lens_all = [np.full(per, min(CHUNK, n_samples - c * CHUNK), np.int32) for c in range(n_chunks)]
    fin_all = [np.full(per, 1 if c == n_chunks - 1 else 0, np.int32) for c in range(n_chunks)]
    stats = {"steps": 0, "launches": 0, "blocks": 0}
    stats_lock = threading.Lock()

    def run_shards(sgroup, glist, fn):
        if sgroup is None:
            return [fn(0, glist[0], 0, S)]
        return sgroup.run_pass(fn)

    def shard_pass_resident(i, g, lo, hi, res=None):
        res = resident if res is None else res
        g.reset()
        loc = [0, 0, 0]
        for c in range(n_chunks):
            st = g.push_device(ids, res[lo:hi], lens_all[c], fin_all[c], col_offset=c * CHUNK)
            loc[0] += st.n_decode_steps; loc[1] += st.n_kernel_launches; loc[2] += st.n_encoder_blocks
        with stats_lock:
            stats["steps"] = max(stats["steps"], 0) + (loc[0] if i == 0 else 0)
            stats["launches"] += loc[1]; stats["blocks"] += loc[2]

    def one_pass_resident():
        run_shards(sg, groups, shard_pass_resident)

    lat_ms = []

    def shard_pass_e2e(i, g, lo, hi):
        """Public API with host buffers: per-chunk H2D from pinned memory inside, results read back."""
        g.reset()
        lat = []
        for c in range(n_chunks):
            t1 = time.perf_counter()
            g.push_batch(ids, host_chunks[c][lo:hi], lens_all[c], fin_all[c])
            lat.append(1000.0 * (time.perf_counter() - t1))
        out = g.results_all(True, True)                # one bulk D2H of every stream's beam
        with stats_lock:
            lat_ms.extend(lat)
        return out

    def one_pass_e2e():
        return [r for part in run_shards(sg, groups, shard_pass_e2e) for r in part]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def final_beams(glist):
        """(yseq per hyp, scores, xpos per hyp, process_idx) of every stream of this rank after a finished pass."""
        torch.cuda.synchronize()
        return [b for g_ in glist for b in g_.beams_all()]

    def common_prefix(a, b):
        n = 0
        for x, y in zip(a, b):
            if x != y:
                break
            n += 1
        return n

    for _ in range(args.warmup):
        one_pass_resident()
    # one extra untimed pass with every kernel wrapped in CUDA events (decode steps sampled 1 in 8): the share of
    # each kernel in the step, used to pick the dominant kernel for the roofline object
    breakdown = None
    if args.profile_kernel == "auto" or args.breakdown:
        grp.profile_begin("all", max_launches=400000, stride=8)       # = STRIDE below
        one_pass_resident()
        raw = grp.profile_end("all")
        raw.pop("_counters", None)
        tot = {k: v for k, v in raw.items() if k in ("decode_step_total", "encoder_total")}
        # decode-step kernels are bracketed on every 8th search iteration only, per-push kernels on every push: scale the
        # sampled figures up so that launches / ms / share describe the whole pass
        STRIDE = 8
        per_step = ("dec_", "ctc_prefix", "ctc_state_update", "prebeam", "combine_topk", "beam_prune", "step_finish")
        parts = {k: ((v[0] * STRIDE, v[1] * STRIDE) if k.startswith(per_step) else v) for k, v in raw.items() if k not in tot}
        s_ms = sum(v[1] for v in parts.values())
        breakdown = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / s_ms, 4)}
                     for k, v in sorted(parts.items(), key=lambda kv: -kv[1][1])}
        breakdown["_totals"] = {k: {"launches": v[0] * (STRIDE if k == "decode_step_total" else 1),
                                    "ms": round(v[1] * (STRIDE if k == "decode_step_total" else 1), 3)} for k, v in tot.items()}
        breakdown["_note"] = ("CUDA-event pairs around every launch of shard 0 in one extra pass; decode-step kernels sampled "
                              "every 8th iteration and scaled x8; bracketing adds ~5 us to each small kernel")
        if args.profile_kernel == "auto":
            rankable = [k for k in parts if k in ("ctc_prefix", "dec_self_attn", "dec_cross_attn", "dec_ffn1", "dec_ffn2",
                                                  "enc_ffn1", "enc_ffn2", "conv2")]
            args.profile_kernel = max(rankable, key=lambda k: parts[k][1])
    # With CUDA-graph replay (--graph) kernels inside a graph cannot be bracketed by events, and a profiled engine
    # falls back to plain launches: the timed region then runs unprofiled and the dominant kernel is timed in a
    # separate pass afterwards.  Without --graph the kernel is timed live inside the timed region (default).
    prof_in_timed = not args.graph
    prof = grp.profile_begin(args.profile_kernel) if prof_in_timed else None
    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = {"steps": 0, "launches": 0, "blocks": 0}
    e0.record()
    for _ in range(args.steps):
        one_pass_resident()
    torch.cuda.synchronize()          # shards run on their own streams: all of them must be done before the stop event
    e1.record()
    sync_all()
    timed_stats = dict(stats)         # later passes (e2e, fp32 mode) must not leak into the timed region's counts
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if prof_in_timed:
        roof = grp.profile_end(prof)
    else:
        prof = grp.profile_begin(args.profile_kernel)
        one_pass_resident()
        torch.cuda.synchronize()
        roof = grp.profile_end(prof)
        roof["timed_in"] = "separate pass after the timed region (graph replay active in the timed region)"
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = world * S * args.seconds * args.steps
    value = audio_s / (ms / 1000.0)

    e2e = None
    if not args.no_e2e:
        one_pass_e2e()                                 # warm the host path
        lat_ms.clear()
        sync_all()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        d2h = 0
        for _ in range(n_e2e):
            out = one_pass_e2e()
            # bytes actually read back: the bulk copies of control words, yseq, xpos and scores of every shard
            d2h = sum(t.numel() * t.element_size() for g_ in groups for t in g_._bulk[1:])
            assert len(out) == S and all(len(r) == args.beam for r in out)
        sync_all()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * S * args.seconds * n_e2e / dt, "unit": "audio-s/s",
               "h2d_bytes_per_step": int(S * n_samples * 4), "d2h_bytes_per_step": int(d2h),
               "p50_chunk_ms": float(statistics.median(lat_ms)), "p95_chunk_ms": float(np.percentile(lat_ms, 95))}

        if world > 1:
            # the one exchange of a sharded run: finished transcripts of every rank's streams as
            # fixed-width records over NCCL / NVLink (outside the timed region; a few MB per rank)
            from speechcatcher_b200.sharding import gather_beams
            tg = time.perf_counter()
            try:
                recs = [[], [], [], []]
                for g_ in groups:
                    ctl_n, ys_n, xp_n, sc_n = g_._read_all()
                    cur, idx = ctl_n[:, 0], np.arange(ctl_n.shape[0])
                    for lst, arr in zip(recs, (ctl_n[:, 1:3], ys_n[cur, idx], xp_n[cur, idx], sc_n[cur, idx])):
                        lst.append(np.ascontiguousarray(arr))
                g_ctl, g_ys, g_xp, g_sc = gather_beams(*[np.concatenate(x) for x in recs], device=dev)
                assert g_ctl.shape[0] == world * S and (g_ctl[:, 0] == args.beam).all()
                e2e["transcript_gather"] = {"streams": int(g_ctl.shape[0]), "ms": 1000.0 * (time.perf_counter() - tg),
                                            "bytes": int(g_ctl.nbytes + g_ys.nbytes + g_xp.nbytes + g_sc.nbytes)}
            except Exception as ex:      # reported in the line, never silently dropped; the throughput numbers stand
                e2e["transcript_gather"] = {"error": f"{type(ex).__name__}: {ex}"}

    # per-kernel rooflines without cross-shard interference: one dedicated group holding all S streams, one full pass
    # per kernel with CUDA-event pairs around every launch of that kernel (same workload, same deferred scheduling)
    extra_roof = None
    sync_latency = None
    if not args.no_extra_rooflines and rank == 0:
        g1 = StreamGroup(md, n_streams=S, beam_size=args.beam, ctc_weight=0.3, device=dev, dtype=args.dtype, use_bbd=False,
                         max_chunk=CHUNK, max_seconds=args.seconds + 1.0, own_stream=True)
        g1.set_option("lazy_threshold", max(1, S - S // 32))
        ids1 = np.arange(S, dtype=np.int32)
        lens1 = [np.full(S, int(l[0]), np.int32) for l in lens_all]
        fin1 = [np.full(S, int(f[0]), np.int32) for f in fin_all]

        def pass1():
            g1.reset()
            blocks = 0
            for c in range(n_chunks):
                blocks += g1.push_device(ids1, resident, lens1[c], fin1[c], col_offset=c * CHUNK).n_encoder_blocks
            return blocks

        pass1()
        # one pass with CUDA-event pairs around EVERY kernel launch (no sampling): launches and time per kernel
        g1.profile_begin("all", max_launches=600000, stride=1)
        n_blocks1 = pass1()
        torch.cuda.synchronize()
        raw1 = g1.profile_end("all")
        cnt1 = raw1.pop("_counters")
        Dm, Fm, Le = g1.cfg.d_model, g1.cfg.ffn, g1.cfg.enc_layers
        enc_flops = {"enc_ffn1": 2.0 * 41 * n_blocks1 * Le * Fm * Dm, "enc_ffn2": 2.0 * 41 * n_blocks1 * Le * Fm * Dm,
                     "enc_qkv": 2.0 * 41 * n_blocks1 * Le * 3 * Dm * Dm, "enc_o": 2.0 * 41 * n_blocks1 * Le * Dm * Dm}
        extra_roof = []
        for kname in ("ctc_prefix", "dec_cross_attn", "dec_self_attn", "enc_ffn1", "enc_ffn2", "enc_qkv", "dec_ffn1"):
            if kname in raw1 and raw1[kname][0] > 0:   # e.g. enc_ffn1 does not exist as a kernel when the FFN is fused
                extra_roof.append(g1.roofline_of(kname, raw1[kname][0], raw1[kname][1], cnt1, enc_flops.get(kname)))
        # synchronous per-chunk latency (what BASELINE's metric puts beside the CPU's per-call latency): strict mode
        # (every push fully decodes its blocks like the reference), host buffers in, results of EVERY chunk read back
        g1.set_option("lazy_threshold", 0)
        g1.reset()
        lat1 = []
        for c in range(n_chunks):
            t1 = time.perf_counter()
            g1.push_batch(ids1, host_chunks[c], lens1[c], fin1[c])
            fin_c = bool(fin1[c][0])
            g1.results_all(fin_c, fin_c)
            lat1.append(1000.0 * (time.perf_counter() - t1))
        sync_latency = {"p50_chunk_ms": float(statistics.median(lat1)), "p95_chunk_ms": float(np.percentile(lat1, 95)),
                        "max_chunk_ms": float(max(lat1)), "streams_per_call": S,
                        "what": "wall time of one synchronous call for all streams of the GPU: pinned host chunk in, every "
                                "block decoded (strict mode), beams of all streams read back"}
        g1.close()
        del g1
        torch.cuda.empty_cache()

    # the CUDA-core fp32 mode on the same workload: its throughput, and the agreement of the benched mode's final
    # beams with it on every stream (`parity`)
    fp32_mode = None
    parity = None
    if args.dtype not in ("float32_simt",) and not args.no_fp32:
        beams_main = final_beams(groups)
        for g_ in groups:
            g_.close()
        if sg is not None:
            sg.pool.shutdown(wait=True)
        del grp, groups, sg, resident
        torch.cuda.empty_cache()
        sg32, groups32 = make_groups("float32_simt")
        res32 = host.to(dev)

        def pass32():
            run_shards(sg32, groups32, lambda i, g, lo, hi: shard_pass_resident(i, g, lo, hi, res32))

        pass32()
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pass32()
        sync_all()
        a1.record()
        sync_all()
        ms32 = a0.elapsed_time(a1)
        if world > 1:
            t = torch.tensor([ms32], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms32 = float(t.item())
        beams32 = final_beams(groups32)
        for g_ in groups32:
            g_.close()
        fp32_mode = {"value": world * S * args.seconds / (ms32 / 1000.0), "unit": "audio-s/s", "ms_per_step": ms32,
                     "steps": 1, "warmup": 1, "note": "float32_simt: true-fp32 CUDA-core GEMMs, n-best identical to the reference"}
        same1 = sum(1 for a, b in zip(beams_main, beams32) if a[0][:1] == b[0][:1])
        samen = sum(1 for a, b in zip(beams_main, beams32) if a[0] == b[0])
        pre = [common_prefix(a[0][0], b[0][0]) / max(1, len(b[0][0])) for a, b in zip(beams_main, beams32)]
        cnt = [same1, samen, len(beams32)]
        if world > 1:
            t = torch.tensor(cnt + [sum(pre)], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            cnt, pre_sum = [int(v) for v in t[:3].tolist()], float(t[3].item())
        else:
            pre_sum = float(sum(pre))
        parity = {"mode": args.dtype, "against": "float32_simt pass of this run (CUDA-core fp32 GEMMs; n-best identical to "
                  "the reference on every golden)", "streams_compared": cnt[2], "audio_seconds_per_stream": args.seconds,
                  "same_1best_frac": cnt[0] / cnt[2], "same_nbest_frac": cnt[1] / cnt[2],
                  "mean_1best_common_prefix_frac": pre_sum / cnt[2],
                  "tokens_per_1best": float(np.mean([len(b[0][0]) for b in beams32]))}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_s = args.cpu_sample_seconds or cpu_sample_seconds(args.warmup + args.steps)   # = the reference arm's
        cpu = CpuBaseline(md, args.beam, cores)
        # warm-up with the SAME sample: the reference's per-chunk tensor shapes grow with the utterance, and the first pass
        # over new shapes is ~2x slower (primitive / allocator caches) -- a 2 s warm-up left the timed pass cold (2.4 vs
        # 4.8 audio-s/s for the reference arm, which times passes after a full-length warm-up)
        cpu.run(sample_s)
        v, p50 = cpu.run(sample_s)
        cpu.close()
        cpu_base = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": cpu.kind, "sample": cpu.describe(sample_s),
                    "p50_chunk_ms": p50, "cpu_model": cpu_model()}
    if rank == 0:
        line = {"metric": "audio-sec/sec (RTFx)", "value": value, "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.dtype == "bfloat16" else "f32", "data": "synthetic",
                "config": {"workload": workload, "l2": f"inputs larger than L2 ({S * n_chunks * CHUNK * 4 / 1e6:.0f} MB of waveforms per GPU)",
                           "mode": {"float32_tc": "float32_tc: fp32 activations, every Linear a split-fp16 (hi + lo) tcgen05 GEMM "
                                                  "with fp32 accumulation (3 UMMAs per product), fp32 attention / KV caches",
                                    "float32_simt": "float32_simt: true-fp32 CUDA-core GEMMs",
                                    "float32": "float32 (engine default fp32 GEMM)",
                                    "bfloat16": "bfloat16: bf16 tcgen05 GEMMs / attention, bf16 KV caches"}[args.dtype],
                           "shards_per_gpu": G, "cuda_graphs": args.graph,
                           "encoder_sm_partition": groups_enc_sms,
                           "decode_scheduling": ("strict: every push drains its decode blocks" if lazy == 0 else
                                                 f"deferred: a push stops iterating below {lazy} active streams; final calls drain"),
                           "decode_steps_per_pass": timed_stats["steps"] // max(1, args.steps),
                           "encoder_blocks_per_pass": timed_stats["blocks"] // max(1, args.steps)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": timed_stats["launches"],
                "roofline": roof, "parity": parity, "sync_latency": sync_latency, "roofline_single_group": extra_roof,
                "cpu_baseline": cpu_base,
                "fp32_mode": fp32_mode,
                "kernel_breakdown_sampled": breakdown}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
