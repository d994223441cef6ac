"""Oracle: one-step Transformer decoder with the ESPnet per-layer output cache.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  speechcatcher/model/decoder/transformer_decoder.py:210-312 (forward_one_step, batch_score)
  speechcatcher/model/decoder/decoder_layer.py:60-132
The cache of layer l holds that layer's outputs for positions 0..L-1; keys and
values are re-projected from it every step exactly as the reference does, so the
CPU-baseline timing of this port reflects the reference's algorithmic cost.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .encoder import ffn, layer_norm, mha, sinusoid_table


class DecoderOracle:
    def __init__(self, W: Dict[str, torch.Tensor], n_layers: int, n_head: int, d_model: int = 256):
        self.W, self.L, self.H, self.D = W, n_layers, n_head, d_model
        self.pe = sinusoid_table(5000, d_model)

    def _layer(self, l, tgt, memory, cache):
        """decoder_layer.py:80-132 (pre-norm, no concat)."""
        W, p = self.W, f"decoder.decoders.{l}"
        residual = tgt
        t = layer_norm(tgt, W[p + ".norm1.weight"], W[p + ".norm1.bias"])
        if cache is None:
            q = t
            # causal mask only matters when more than one position is processed at once
            n = t.size(1)
            mask = torch.tril(torch.ones(n, n, dtype=torch.bool)).unsqueeze(0)
        else:
            q = t[:, -1:, :]
            residual = residual[:, -1:, :]
            mask = None
        x = residual + mha(W, p + ".self_attn", self.H, q, t, t, mask)
        n2 = layer_norm(x, W[p + ".norm2.weight"], W[p + ".norm2.bias"])
        x = x + mha(W, p + ".src_attn", self.H, n2, memory, memory, None)
        n3 = layer_norm(x, W[p + ".norm3.weight"], W[p + ".norm3.bias"])
        x = x + ffn(W, p + ".feed_forward", n3)
        if cache is not None:
            x = torch.cat([cache, x], dim=1)
        return x

    def batch_score(self, ys: torch.Tensor, states: List[Optional[List[torch.Tensor]]],
                    memory: torch.Tensor) -> Tuple[torch.Tensor, List[List[torch.Tensor]]]:
        """ys (n, L) int64; states per hyp = list over layers of (L-1, D); memory (n, T, D).
        Returns log-probs (n, V) and the new per-hyp caches (transformer_decoder.py:275-312)."""
        n = ys.size(0)
        if states[0] is None:
            cache = [None] * self.L
        else:
            cache = [torch.stack([states[b][i] for b in range(n)]) for i in range(self.L)]
        W = self.W
        x = F.embedding(ys, W["decoder.embed.0.weight"]) * math.sqrt(self.D) + self.pe[: ys.size(1)].unsqueeze(0)
        new_cache = []
        for l in range(self.L):
            x = self._layer(l, x, memory, cache[l])
            new_cache.append(x)
        y = layer_norm(x[:, -1], W["decoder.after_norm.weight"], W["decoder.after_norm.bias"])
        logp = torch.log_softmax(F.linear(y, W["decoder.output_layer.weight"], W["decoder.output_layer.bias"]), dim=-1)
        return logp, [[new_cache[i][b] for i in range(self.L)] for b in range(n)]
