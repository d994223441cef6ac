"""ShardedStreamGroup: several StreamGroups on one GPU, each on its own CUDA stream and host thread.

The search is a chain of ~165 small dependent kernels per iteration, i.e. latency-bound: one chain cannot fill
148 SMs no matter how it is tuned.  Streams are independent, so the S streams of a GPU are split into `n_shards`
groups whose chains run concurrently (and no longer wait for each other at push boundaries).  The C ABI releases
the GIL (ctypes), so one Python thread per shard drives its engine.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import List, Sequence

import numpy as np
import torch

from .stream_group import StreamGroup


class ShardedStreamGroup:
    def __init__(self, model_dir, n_streams: int, n_shards: int = 2, device: str = "cuda:0", **kw):
        assert n_streams % n_shards == 0, "n_streams must be a multiple of n_shards"
        self.device = torch.device(device)
        self.n_streams, self.n_shards, self.per = n_streams, n_shards, n_streams // n_shards
        # high priority: the search chains of small kernels go ahead of the engines' (low-priority) encoder streams
        self.cuda_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(n_shards)]
        self.shards: List[StreamGroup] = []
        for st in self.cuda_streams:                    # built one after another: shared constant tables
            with torch.cuda.stream(st):
                self.shards.append(StreamGroup(model_dir, n_streams=self.per, device=device, **kw))
        torch.cuda.synchronize(self.device)
        self.pool = ThreadPoolExecutor(max_workers=n_shards)
        self.beam_size = self.shards[0].beam_size

    def close(self):
        self.pool.shutdown(wait=True)
        for g in self.shards:
            g.close()

    def set_option(self, name: str, value: int):
        for g in self.shards:
            g.set_option(name, value)

    def shard_of(self, stream: int):
        return self.shards[stream // self.per], stream % self.per

    def _each(self, fn):
        """Run fn(shard_index, group) on every shard's thread and CUDA stream; re-raise the first failure."""
        def run(i):
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self.cuda_streams[i]):
                return fn(i, self.shards[i])
        return [f.result() for f in [self.pool.submit(run, i) for i in range(self.n_shards)]]

    def reset(self):
        self._each(lambda i, g: g.reset())

    def run_pass(self, fn):
        """fn(shard_index, group, lo, hi): the caller's per-shard work for global streams [lo, hi)."""
        return self._each(lambda i, g: fn(i, g, i * self.per, (i + 1) * self.per))

    def synchronize(self):
        for st in self.cuda_streams:
            st.synchronize()

    def beam(self, stream: int):
        g, s = self.shard_of(stream)
        return g.beam(s)

    def results(self, stream: int, is_final: bool, finalize_all: bool, token_list=None):
        g, s = self.shard_of(stream)
        return g.results(s, is_final, finalize_all, token_list)
