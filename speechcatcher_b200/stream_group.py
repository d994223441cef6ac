"""StreamGroup: the batched, device-resident engine under the Speech2TextStreaming facade.

One StreamGroup per GPU holds S independent streams; `push` advances any subset of them by one
chunk each (one reference `Speech2TextStreaming.__call__` per stream,
speechcatcher/speech2text_streaming.py:402-539) in a single batched pass through the CUDA path.
PyTorch is used for device memory, streams and H2D copies only.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import yaml

from . import _lib
from ._lib import ScConfig, ScPushStats, ScStreamPlan
from .model_files import find_checkpoint, find_stats, read_stats, state_dict_of
from .weights import bf16_names, pack_weights, split_f16

# engine precision codes (include/speechcatcher_b200.h ScConfig.precision)
PRECISIONS = {"float32_simt": 0, "bfloat16": 1, "float32_tc": 2}
DEFAULT_FP32 = "float32_tc"       # what dtype="float32" means (both pass the reference goldens exactly)

EOS_FILTER_ID = 1023   # hard-coded in the reference's output filter (speech2text_streaming.py:474)


def _find_checkpoint(model_dir: Path) -> Path:
    return find_checkpoint(model_dir)


def load_stats(path: Path):
    return read_stats(path)


def _frontend_tables():
    """hann(400) and the slaney mel matrix exactly as the reference builds them
    (model/frontend/stft_frontend.py:68-85: torch.hann_window + torchaudio melscale_fbanks)."""
    import torchaudio
    window = torch.hann_window(400)
    mel = torchaudio.functional.melscale_fbanks(n_freqs=257, f_min=0.0, f_max=8000.0, n_mels=80,
                                                sample_rate=16000, norm="slaney", mel_scale="slaney")
    return window.contiguous(), mel.contiguous()


class StreamGroup:
    def __init__(self, model_dir, n_streams: int = 1, beam_size: int = 5, ctc_weight: float = 0.3,
                 device: str = "cuda:0", dtype: str = "float32", use_bbd: bool = False,
                 max_chunk: int = 8192, max_seconds: float = 61.0, own_stream: bool = False):
        if not str(device).startswith("cuda"):
            raise RuntimeError("speechcatcher_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: the B200 path cannot run")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.model_dir = Path(model_dir)
        self.n_streams, self.beam_size, self.ctc_weight, self.use_bbd = n_streams, beam_size, ctc_weight, use_bbd
        import os
        if dtype == "float32":          # which GEMM kernel serves the fp32 mode: SCB_FP32_GEMM = "tc" | "simt"
            dtype = {"tc": "float32_tc", "simt": "float32_simt"}.get(os.environ.get("SCB_FP32_GEMM", ""), DEFAULT_FP32)
        if dtype not in PRECISIONS:
            raise ValueError("dtype must be 'float32' (parity mode; 'float32_tc' = split-fp16 tensor-core GEMMs, "
                             "'float32_simt' = CUDA-core GEMMs) or 'bfloat16' (bf16 tensor-core mode)")
        self.dtype = dtype
        self.precision = PRECISIONS[dtype]
        ckpt = torch.load(_find_checkpoint(self.model_dir), map_location="cpu")
        sd = state_dict_of(ckpt)
        vocab = sd["decoder.embed.0.weight"].shape[0]
        cfg_path = self.model_dir / "config.yaml"
        conf = yaml.safe_load(open(cfg_path)) if cfg_path.exists() else {}
        enc, dec = conf.get("encoder_conf", {}), conf.get("decoder_conf", {})
        self.max_chunk = int(max_chunk)
        self.max_seconds = float(max_seconds)
        max_frames = self.check_capacity(max_seconds, max_chunk)
        self.cfg = ScConfig(
            d_model=enc.get("output_size", 256), enc_heads=enc.get("attention_heads", 4),
            enc_layers=enc.get("num_blocks", 12), dec_heads=dec.get("attention_heads", 4),
            dec_layers=dec.get("num_blocks", 6), vocab=vocab, ffn=2048, n_streams=n_streams, beam=beam_size,
            max_chunk=self.max_chunk, max_frames=max_frames, use_bbd=int(use_bbd), precision=self.precision,
            ctc_weight=ctc_weight)
        with torch.cuda.device(self.device):
            need = C.c_size_t()
            _lib.check(self.lib.sc_engine_workspace_bytes(C.byref(self.cfg), C.byref(need)), "workspace_bytes")
            self.workspace = torch.zeros(need.value + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            aligned = (base + 255) // 256 * 256
            self.handle = C.c_void_p()
            _lib.check(self.lib.sc_engine_create(C.byref(self.cfg), C.c_void_p(aligned), need.value,
                                                 C.byref(self.handle)), "create")
            packed = pack_weights(sd, self.cfg.enc_layers, self.cfg.dec_layers, self.cfg.d_model)
            self.weights: Dict[str, torch.Tensor] = {}
            for k, v in packed.items():
                self.weights[k] = v.to(self.device)
            if self.precision == 1:
                for k in bf16_names(self.cfg.enc_layers, self.cfg.dec_layers):
                    self.weights[k + ".bf16"] = self.weights[k].to(torch.bfloat16).contiguous()
            if self.precision == 2:
                for k in bf16_names(self.cfg.enc_layers, self.cfg.dec_layers):
                    self.weights[k + ".x3"] = split_f16(self.weights[k])
            for k, v in self.weights.items():
                _lib.check(self.lib.sc_engine_set_weight(self.handle, k.encode(), C.c_void_p(v.data_ptr()),
                                                         v.numel()), f"set_weight({k})")
            window, mel = _frontend_tables()
            self.mean, self.std = find_stats(self.model_dir)
            mean_p = self.mean.ctypes.data_as(C.c_void_p) if self.mean is not None else None
            std_p = self.std.ctypes.data_as(C.c_void_p) if self.std is not None else None
            _lib.check(self.lib.sc_engine_set_frontend(self.handle, C.c_void_p(window.data_ptr()),
                                                       C.c_void_p(mel.data_ptr()), mean_p, std_p), "set_frontend")
            _lib.check(self.lib.sc_engine_finalize(self.handle), "finalize")
        # The engine runs on the stream that is current at construction.  own_stream=True gives it a stream of its own
        # (ordered after the caller's stream at every push): needed for the CUDA-graph replay when the caller works on
        # the legacy default stream, which cannot be captured.
        own_stream = own_stream or os.environ.get("SCB_OWN_STREAM") == "1"     # switch for whole-suite validation runs
        self.own_stream = bool(own_stream)
        # the search is a chain of small dependent kernels: its stream gets the high priority so that its CTAs are
        # scheduled ahead of the (large) encoder kernels of the engine's second stream
        self.stream = (torch.cuda.Stream(device=self.device, priority=-1) if own_stream
                       else torch.cuda.current_stream(self.device))
        self._wave_dev = torch.zeros(n_streams, self.max_chunk, dtype=torch.float32, device=self.device)
        self._wave_host = torch.zeros(n_streams, self.max_chunk, dtype=torch.float32).pin_memory()
        self._h2d_done = None
        self.last_stats = ScPushStats()
        self.total_launches = 0

    @staticmethod
    def check_capacity(max_seconds: float, max_chunk: int) -> int:
        """Validates the per-stream capacities an engine would be created with (before anything is allocated) and
        returns the encoder-frame capacity."""
        if int(max_chunk) < 1:
            raise ValueError(f"max_chunk={max_chunk} must be positive")
        max_frames = int(max_seconds * 25) + 64
        if max_frames + 64 > 5000:
            raise ValueError(f"max_seconds={max_seconds:g} exceeds what one un-reset stream can hold: the model's positional "
                             "table has 5000 positions (reference positional_encoding.py:31), i.e. at most 194 s")
        return max_frames

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "handle", None):
            self.lib.sc_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        """Engine options: "lazy_threshold" (deferred decoding, see include/speechcatcher_b200.h), "mma_attention"."""
        _lib.check(self.lib.sc_engine_set_option(self.handle, name.encode(), int(value)), f"set_option({name})")

    def reset(self, streams: Optional[Sequence[int]] = None):
        ids = np.asarray(list(range(self.n_streams)) if streams is None else list(streams), np.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sc_engine_reset(self.handle, ids.ctypes.data_as(C.c_void_p), len(ids),
                                                C.c_void_p(self.stream.cuda_stream)), "reset")

    # ------------------------------------------------------------------ the hot call
    def push(self, streams: Sequence[int], chunks: Sequence[np.ndarray], is_final: Sequence[bool]) -> ScPushStats:
        """One chunk per listed stream (host buffers; the H2D copy is part of the call)."""
        n = len(streams)
        ids = np.asarray(streams, np.int32)
        lens = np.asarray([len(c) for c in chunks], np.int32)
        fin = np.asarray([1 if f else 0 for f in is_final], np.int32)
        if n and lens.max() > self.max_chunk:
            raise ValueError(f"chunk of {int(lens.max())} samples exceeds max_chunk={self.max_chunk}")
        host = self._wave_host
        # the pinned staging buffer is reused: in deferred / overlapped mode a push can return before its H2D copy has
        # drained, so wait for the previous copy before overwriting the buffer (almost always already complete)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        for i, (s, c) in enumerate(zip(ids, chunks)):
            if len(c):
                host[s, : len(c)] = torch.as_tensor(np.asarray(c, np.float32))
        with torch.cuda.device(self.device):
            # own_stream: the copy goes on the engine's stream so that it is ordered after the previous push's reads
            copy_stream = self.stream if self.own_stream else torch.cuda.current_stream(self.device)
            with torch.cuda.stream(copy_stream):
                self._wave_dev.copy_(host, non_blocking=True)
            self._h2d_done = torch.cuda.Event()
            self._h2d_done.record(copy_stream)
        return self.push_device(ids, self._wave_dev, lens, fin)

    def max_feature_frames(self) -> int:
        """Feature frames one stream may push per call (sc_engine_push_features)."""
        return (self.max_chunk + 400) // 160 + 2      # = the STFT frames of the largest waveform slab (make_caps)

    def push_features(self, streams: Sequence[int], feats: Sequence[np.ndarray], is_final: Sequence[bool]) -> ScPushStats:
        """Pre-computed, already normalised feature frames ([n_i, 80] float32 per listed stream) instead of waveforms:
        the 2-D / 3-D input mode of the reference's __call__ (speech2text_streaming.py:438-450)."""
        n = len(streams)
        ids = np.ascontiguousarray(streams, np.int32)
        counts = np.asarray([len(f) for f in feats], np.int32)
        fin = np.asarray([1 if f else 0 for f in is_final], np.int32)
        if n and counts.max() > self.max_feature_frames():
            raise ValueError(f"{int(counts.max())} feature frames exceed the per-push capacity "
                             f"{self.max_feature_frames()} of this engine; construct with a larger max_chunk")
        ld = max(1, int(counts.max()) if n else 1) * 80
        host = torch.zeros(max(n, 1), ld, dtype=torch.float32)
        for i, f in enumerate(feats):
            f = np.ascontiguousarray(f, np.float32)
            if f.ndim != 2 or f.shape[1] != 80:
                raise ValueError(f"features must be [frames, 80], got {f.shape}")
            if len(f):
                host[i, : f.size] = torch.from_numpy(f.reshape(-1))
        st = ScPushStats()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._feat_dev = host.to(self.device)          # kept alive until the next call (the copy is asynchronous)
            rc = self.lib.sc_engine_push_features(self.handle, C.c_void_p(self._feat_dev.data_ptr()), ld,
                                                  ids.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                                  fin.ctypes.data_as(C.c_void_p), n,
                                                  C.c_void_p(self.stream.cuda_stream), C.byref(st))
        _lib.check(rc, "push_features")
        self.last_stats = st
        self.total_launches += st.n_kernel_launches
        return st

    def push_batch(self, ids: np.ndarray, wave_host: torch.Tensor, lens: np.ndarray, fin: np.ndarray) -> ScPushStats:
        """Host waveforms as one [n_streams, L] float32 tensor (row s = stream s); one pinned H2D copy."""
        L = wave_host.shape[1]
        if L > self.max_chunk:
            raise ValueError(f"chunk of {L} samples exceeds max_chunk={self.max_chunk}")
        with torch.cuda.device(self.device):
            # like push(): with an engine-owned stream the copy is issued there, so that it is ordered after the previous
            # push's reads of the staging buffer (a deferred / overlapped push returns before they have drained);
            # otherwise the engine works on the caller's current stream and plain stream order does the same
            copy_stream = self.stream if self.own_stream else torch.cuda.current_stream(self.device)
            with torch.cuda.stream(copy_stream):
                self._wave_dev[:, :L].copy_(wave_host, non_blocking=True)
        return self.push_device(ids, self._wave_dev, lens, fin)

    def push_device(self, ids: np.ndarray, wave_dev: torch.Tensor, lens: np.ndarray, fin: np.ndarray,
                    col_offset: int = 0) -> ScPushStats:
        """Same, with the waveforms already resident: wave_dev[s, col_offset : col_offset+lens[i]] for s = ids[i]."""
        assert wave_dev.dtype == torch.float32 and wave_dev.is_contiguous() and wave_dev.shape[0] == self.n_streams
        ids = np.ascontiguousarray(ids, np.int32)
        lens = np.ascontiguousarray(lens, np.int32)
        fin = np.ascontiguousarray(fin, np.int32)
        st = ScPushStats()
        with torch.cuda.device(self.device):
            if self.own_stream:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))     # the caller's copies into wave_dev
            rc = self.lib.sc_engine_push(self.handle, C.c_void_p(wave_dev.data_ptr() + 4 * col_offset), wave_dev.shape[1],
                                         ids.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p),
                                         fin.ctypes.data_as(C.c_void_p), len(ids),
                                         C.c_void_p(self.stream.cuda_stream), C.byref(st))
        _lib.check(rc, "push")
        self.last_stats = st
        self.total_launches += st.n_kernel_launches
        return st

    # ------------------------------------------------------------------ results
    def beam(self, stream: int) -> Tuple[List[List[int]], List[float], List[List[int]], int]:
        """(yseq per hyp, fp64 scores, xpos per hyp, process_idx) of the running hypotheses."""
        L = 1024
        yseq = np.zeros((self.beam_size, L), np.int32)
        xpos = np.zeros((self.beam_size, L), np.int32)
        score = np.zeros(self.beam_size, np.float64)
        n_hyp, ln, pidx = C.c_int32(), C.c_int32(), C.c_int32()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sc_engine_read_beam(
                self.handle, stream, L, C.byref(n_hyp), C.byref(ln), C.byref(pidx),
                yseq.ctypes.data_as(C.c_void_p), xpos.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p),
                C.c_void_p(self.stream.cuda_stream)), "read_beam")
        n, l = n_hyp.value, ln.value
        return ([yseq[h, :l].tolist() for h in range(n)], score[:n].tolist(),
                [xpos[h, :l].tolist() for h in range(n)], pidx.value)

    # ------------------------------------------------------------------ step trace / counters (parity tests)
    def counter(self, name: str) -> int:
        v = C.c_int64()
        _lib.check(self.lib.sc_engine_counter(self.handle, name.encode(), C.byref(v)), f"counter({name})")
        return int(v.value)

    def trace_begin(self, max_steps: int):
        """Record the score tensors of the next `max_steps` search iterations (sc_engine_set_trace)."""
        off = (C.c_int64 * 8)()
        rec = C.c_int64()
        _lib.check(self.lib.sc_engine_trace_layout(self.handle, off, C.byref(rec)), "trace_layout")
        self._trace_off, self._trace_rec = [int(o) for o in off], int(rec.value)
        self._trace_dev = torch.zeros(max(1, max_steps) * self._trace_rec, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.sc_engine_set_trace(self.handle, C.c_void_p(self._trace_dev.data_ptr()),
                                                self._trace_dev.numel(), max_steps), "set_trace")

    def trace_end(self) -> List[dict]:
        """One dict per recorded iteration with the active rows only: rows (stream, hyp), logp [n, V], pre_ids [n, 40],
        psi [n, 40], psi_eos [n], s_prev [n] (CTC prefix score of each row's hypothesis)."""
        n_steps = self.counter("trace_steps")
        torch.cuda.synchronize(self.device)
        raw = self._trace_dev.cpu().numpy()
        _lib.check(self.lib.sc_engine_set_trace(self.handle, None, 0, 0), "set_trace(off)")
        S, B, V, R = self.n_streams, self.beam_size, self.cfg.vocab, self.n_streams * self.beam_size
        o = self._trace_off
        out = []
        for i in range(n_steps):
            r = raw[i * self._trace_rec:(i + 1) * self._trace_rec]
            f32 = lambda k, n: r[o[k]:o[k] + 4 * n].view(np.float32)
            i32 = lambda k, n: r[o[k]:o[k] + 4 * n].view(np.int32)
            n = int(i32(0, 1)[0])
            row_sh = i32(1, R)[:n].copy()
            ctl = i32(7, S * 16).reshape(S, 16)
            ctc_s = f32(6, 2 * S * B).reshape(2, S, B)
            st, hy = row_sh // B, row_sh % B
            out.append(dict(rows=list(zip(st.tolist(), hy.tolist())), logp=f32(2, R * V).reshape(R, V)[:n].copy(),
                            pre_ids=i32(3, R * 40).reshape(R, 40)[:n].copy(), psi=f32(4, R * 40).reshape(R, 40)[:n].copy(),
                            psi_eos=f32(5, R)[:n].copy(), s_prev=ctc_s[ctl[st, 0], st, hy].copy(),
                            Tb=ctl[st, 7].copy(), length=ctl[st, 2].copy()))
        return out

    # ------------------------------------------------------------------ live kernel timing (bench.py roofline)
    PROF_TAGS = {"ctc_prefix": 1, "dec_self_attn": 2, "dec_cross_attn": 3, "dec_ffn1": 4, "enc_ffn1": 5, "prebeam": 6,
                 "enc_attn": 7, "conv2": 8, "dec_ffn2": 9, "enc_ffn2": 10, "ctc_state_update": 11, "frontend": 12,
                 "conv1": 13, "sub_out": 14, "block_assemble": 15, "enc_ln": 16, "enc_qkv": 17, "enc_o": 18,
                 "enc_handover": 19, "stitch_norm": 20, "ctc_head": 21, "cross_kv": 22, "dec_embed": 23, "dec_ln": 24,
                 "dec_qkv": 25, "dec_self_o": 26, "dec_cross_q": 27, "dec_cross_o": 28, "dec_out": 29,
                 "combine_topk": 30, "beam_prune": 31, "step_finish": 32, "decode_step_total": 33, "encoder_total": 34}
    N_TAGS = 35

    def profile_begin(self, kernel: str = "all", max_launches: int = 60000, stride: int = 1) -> str:
        tag = -1 if kernel == "all" else self.PROF_TAGS[kernel]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sc_engine_profile_begin(self.handle, tag, max_launches, stride), "profile_begin")
        return kernel

    def profile_end(self, kernel: str) -> dict:
        """kernel == "all": {name: (launches, ms)} over the sampled launches.  Otherwise the roofline object for
        bench.py: algorithmic bytes (or FLOPs) of the tagged kernel's launches / their CUDA-event time, against the
        measured peak in MEASURED_PEAKS.json (else the profiling recipe's fallback)."""
        import json
        n = (C.c_int32 * self.N_TAGS)()
        ms = (C.c_double * self.N_TAGS)()
        fl = C.c_double()
        cnt = (C.c_uint64 * 8)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sc_engine_profile_end(self.handle, self.N_TAGS, n, ms, C.byref(fl), cnt), "profile_end")
        if kernel == "all":
            out = {name: (int(n[t]), float(ms[t])) for name, t in self.PROF_TAGS.items() if n[t]}
            out["_counters"] = [int(x) for x in cnt]
            return out
        t = self.PROF_TAGS[kernel]
        return self.roofline_of(kernel, int(n[t]), float(ms[t]), [int(x) for x in cnt], fl.value)

    def roofline_of(self, kernel: str, launches: int, ms: float, cnt, flops: float = None) -> dict:
        """Roofline object of one kernel from its launches / CUDA-event time and the device counters of the same pass:
        algorithmic bytes (or FLOPs) / time against the measured peak in MEASURED_PEAKS.json (else the profiling
        recipe's fallback).  `traffic` = DRAM bytes per launch of that kernel from the committed ncu capture
        (profiles/r2_ncu_traffic.json), null when none exists."""
        import json
        root = Path(__file__).resolve().parent.parent
        peaks, src = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}, "fallback"
        if (root / "MEASURED_PEAKS.json").exists():
            peaks, src = json.load(open(root / "MEASURED_PEAKS.json")), "measured"
        D, F = self.cfg.d_model, self.cfg.ffn
        sec = max(ms, 1e-9) / 1e3
        hbm = {"ctc_prefix": cnt[0], "dec_cross_attn": cnt[2], "dec_self_attn": cnt[3]}
        if kernel in hbm:
            nbytes, extra = hbm[kernel], {}
            if kernel == "dec_self_attn" and cnt[5] > 0:
                # the SURVEY formula (2 L D e bytes per HYPOTHESIS and layer) charges K|V once per hypothesis; the kernel reads
                # the shared ancestor chain once per STREAM, so the roofline is stated against the bytes it moves
                extra = {"bytes": "moved (key-list rows x 4 planes x d_k fp16 per stream, head and layer)",
                         "survey_formula_GBs": cnt[3] / sec / 1e9}
                nbytes = cnt[5]
            ach = nbytes / sec / 1e9
            peak = float(peaks["hbm_gbs"])
            out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, **extra}
        else:
            if flops is None or flops <= 0:
                # decoder FFN GEMMs: 2 F D per active row and LAYER (cnt[1] counts active rows once per search iteration)
                flops = 2.0 * cnt[1] * F * D * self.cfg.dec_layers
            ach = flops / sec / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
            out = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak}
            if self.precision == 2:
                out["note"] = ("algorithmic FLOPs (2 M N K, 41 useful rows per encoder block) of a split-fp16 GEMM: the "
                               "tensor cores execute 3 UMMAs per product, so frac <= 1/3 by construction; "
                               "frac_of_3x_bound = 3 * frac")
                out["frac_of_3x_bound"] = 3.0 * ach / peak
        traffic = None
        tpath = root / "profiles" / "r2_ncu_traffic.json"
        if tpath.exists():
            rec = json.load(open(tpath)).get(self.dtype, {}).get(kernel)
            if rec:
                traffic = rec
        out.update({"kernel": kernel, "launches": launches, "kernel_ms_total": ms, "peak_source": src,
                    "traffic": traffic, "search_iterations": int(cnt[4]), "active_rows_total": int(cnt[1])})
        return out

    def _read_all(self):
        """Bulk D2H of control words, yseq, xpos and scores of every stream into pinned buffers (numpy views)."""
        S, B = self.n_streams, self.beam_size
        if not hasattr(self, "_bulk"):
            lcap = C.c_int32()
            _lib.check(self.lib.sc_engine_token_capacity(self.handle, C.byref(lcap)), "token_capacity")
            L = lcap.value
            self._bulk = (L, torch.empty(S, 16, dtype=torch.int32).pin_memory(),
                          torch.empty(2, S, B, L, dtype=torch.int32).pin_memory(),
                          torch.empty(2, S, B, L, dtype=torch.int32).pin_memory(),
                          torch.empty(2, S, B, dtype=torch.float64).pin_memory())
        L, ctl, ys, xp, sc = self._bulk
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sc_engine_read_all(self.handle, C.c_void_p(ctl.data_ptr()), C.c_void_p(ys.data_ptr()),
                                                   C.c_void_p(xp.data_ptr()), C.c_void_p(sc.data_ptr()),
                                                   C.c_void_p(self.stream.cuda_stream)), "read_all")
        return ctl.numpy(), ys.numpy(), xp.numpy(), sc.numpy()

    def beams_all(self):
        """Beams of every stream with four bulk D2H copies into pinned buffers:
        list over streams of (yseq per hyp, scores, xpos per hyp, process_idx)."""
        ctl_n, ys_n, xp_n, sc_n = self._read_all()
        out = []
        for s in range(self.n_streams):
            cur, n, ln, pidx = (int(v) for v in ctl_n[s, :4])
            out.append(([ys_n[cur, s, h, :ln].tolist() for h in range(n)], sc_n[cur, s, :n].tolist(),
                        [xp_n[cur, s, h, :ln].tolist() for h in range(n)], pidx))
        return out

    def results_all(self, is_final: bool, finalize_all: bool, token_list=None):
        """`results` for every stream from one bulk read-back (hypotheses stay numpy rows until the final lists)."""
        ctl_n, ys_n, xp_n, sc_n = self._read_all()
        out = []
        for s in range(self.n_streams):
            cur, n, ln, pidx = (int(v) for v in ctl_n[s, :4])
            out.append(self._assemble((ys_n[cur, s, :n, :ln], sc_n[cur, s, :n].tolist(), xp_n[cur, s, :n, :ln], pidx),
                                      is_final, finalize_all, token_list))
        return out

    def last_plan(self, stream: int) -> ScStreamPlan:
        p = ScStreamPlan()
        _lib.check(self.lib.sc_engine_last_plan(self.handle, stream, C.byref(p)), "last_plan")
        return p

    def buffer(self, name: str, dtype=torch.float32) -> torch.Tensor:
        """Debug view of a named internal device buffer (tests)."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.sc_engine_buffer(self.handle, name.encode(), C.byref(ptr), C.byref(n)), "buffer")
        base = self.workspace.data_ptr()
        off = ptr.value - base
        nbytes = n.value * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off: off + nbytes].view(dtype)

    def results(self, stream: int, is_final: bool, finalize_all: bool, token_list=None):
        """Output assembly of Speech2TextStreaming.__call__ (speech2text_streaming.py:466-539)."""
        return self._assemble(self.beam(stream), is_final, finalize_all, token_list)

    _TOK_TABLES = {}

    @classmethod
    def _tok_table(cls, token_list):
        """Token strings as a numpy object array so a hypothesis is turned into strings with one fancy index."""
        key = id(token_list) if token_list is not None else None
        hit = cls._TOK_TABLES.get(key)
        if hit is None or hit[0] is not token_list:
            names = token_list if token_list is not None else [str(i) for i in range(65536)]
            arr = np.empty(len(names), dtype=object)
            arr[:] = list(names)
            hit = cls._TOK_TABLES[key] = (token_list, arr)
        return hit[1]

    @classmethod
    def _assemble(cls, beam, is_final: bool, finalize_all: bool, token_list=None):
        """Output assembly of Speech2TextStreaming.__call__ (speech2text_streaming.py:466-539), vectorised with numpy
        (a 60 s hypothesis holds ~600 tokens; per-token Python loops would dominate the end-to-end time)."""
        yseqs, scores, xposs, _ = beam
        out = []
        table = cls._tok_table(token_list)
        for y, sc, xp in zip(yseqs, scores, xposs):
            y, xp = np.array(y, dtype=np.int32), np.array(xp, dtype=np.int32)     # owned copies (hyp.yseq / hyp.xpos)
            if (not is_final or not finalize_all) and y[-1] != EOS_FILTER_ID:
                continue
            if is_final:
                ids, pos = y[1:], xp[1:]
                if ids.size and ids[-1] == EOS_FILTER_ID:
                    ids, pos = ids[:-1], pos[:-1]
                keep = (ids > 1) & (ids != EOS_FILTER_ID)     # drops blank 0, sos/eos 1 and the hard-coded 1023
                ids, pos = ids[keep], pos[keep]
            else:
                ids = pos = np.zeros(0, np.int32)  # output_index is always 0 in the reference
            toks = table[ids].tolist()
            if token_list is not None:
                text = "".join(toks).replace("\u2581", " ").strip()
            else:
                text = " ".join(toks)
            out.append((text, toks, ids.tolist(), pos.tolist(), dict(yseq=y, score=sc, xpos=xp)))
        return out

    @classmethod
    def _assemble_rowwise(cls, beam, is_final: bool, finalize_all: bool, token_list=None):
        """One hypothesis at a time, following speech2text_streaming.py:466-539 line by line (the checker of the
        numpy `_assemble` in tests/test_assemble_cpu.py)."""
        yseqs, scores, xposs, _ = beam
        out = []
        for y, sc, xp in zip(yseqs, scores, xposs):
            y, xp = [int(t) for t in y], [int(t) for t in xp]
            if (not is_final or not finalize_all) and y[-1] != EOS_FILTER_ID:
                continue
            if is_final:
                ids, pos = y[1:], xp[1:]
                if ids and ids[-1] == EOS_FILTER_ID:
                    ids, pos = ids[:-1], pos[:-1]
            else:
                ids, pos = [], []
            keep = [i for i, t in enumerate(ids) if t not in (0, 1, EOS_FILTER_ID)]
            ids, pos = [ids[i] for i in keep], [pos[i] for i in keep]
            if token_list is not None:
                toks = [token_list[t] for t in ids]
                text = "".join(toks).replace("\u2581", " ").strip()
            else:
                toks = [str(t) for t in ids]
                text = " ".join(toks)
            out.append((text, toks, ids, pos, dict(yseq=np.array(y, np.int32), score=sc, xpos=np.array(xp, np.int32))))
        return out
