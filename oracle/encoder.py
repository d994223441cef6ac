"""Oracle: conv2d subsampling + contextual-block streaming Transformer encoder.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, for inference only,
  speechcatcher/model/encoder/subsampling.py:71-106
  speechcatcher/model/encoder/contextual_block_transformer_encoder.py:241-419, 500-528
  speechcatcher/model/encoder/contextual_block_encoder_layer.py:178-271
  speechcatcher/model/attention/multi_head_attention.py:63-133   (vanilla path)
  speechcatcher/model/layers/positional_encoding.py:26-75, feed_forward.py:41-50,
  normalization.py:23 (eps 1e-12)
Weights come as a dict keyed like the reference state_dict.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BLOCK, HOP_B, LOOK = 40, 16, 16
LN_EPS = 1e-12


def sinusoid_table(max_len: int, d: int) -> torch.Tensor:
    """positional_encoding.py:38-46."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def layer_norm(x, w, b):
    return F.layer_norm(x, (x.size(-1),), w, b, LN_EPS)


def mha(W: Dict[str, torch.Tensor], p: str, n_head: int, q_in, k_in, v_in, mask=None):
    """Vanilla multi-head attention (multi_head_attention.py:63-133).
    q_in (B, Tq, D), k_in/v_in (B, Tk, D), mask (B, Tq, Tk) or None."""
    B, D = q_in.size(0), q_in.size(-1)
    dk = D // n_head
    q = F.linear(q_in, W[p + ".linear_q.weight"], W[p + ".linear_q.bias"]).view(B, -1, n_head, dk).transpose(1, 2)
    k = F.linear(k_in, W[p + ".linear_k.weight"], W[p + ".linear_k.bias"]).view(B, -1, n_head, dk).transpose(1, 2)
    v = F.linear(v_in, W[p + ".linear_v.weight"], W[p + ".linear_v.bias"]).view(B, -1, n_head, dk).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        m = mask.unsqueeze(1)
        scores = scores.masked_fill(m == 0, torch.finfo(scores.dtype).min)
        attn = torch.softmax(scores, dim=-1).masked_fill(m == 0, 0.0)
    else:
        attn = torch.softmax(scores, dim=-1)
    x = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, D)
    return F.linear(x, W[p + ".linear_out.weight"], W[p + ".linear_out.bias"])


def ffn(W, p, x):
    h = torch.relu(F.linear(x, W[p + ".w_1.weight"], W[p + ".w_1.bias"]))
    return F.linear(h, W[p + ".w_2.weight"], W[p + ".w_2.bias"])


def conv2d_subsample(W, x: torch.Tensor) -> torch.Tensor:
    """(1, T, 80) -> (1, T', D)  (subsampling.py:87-96)."""
    h = x.unsqueeze(1)
    h = torch.relu(F.conv2d(h, W["encoder.embed.conv.0.weight"], W["encoder.embed.conv.0.bias"], stride=2))
    h = torch.relu(F.conv2d(h, W["encoder.embed.conv.2.weight"], W["encoder.embed.conv.2.bias"], stride=2))
    b, c, t, f = h.size()
    h = h.transpose(1, 2).contiguous().view(b, t, c * f)
    return F.linear(h, W["encoder.embed.out.weight"], W["encoder.embed.out.bias"])


class EncoderOracle:
    """Streaming state machine of ContextualBlockTransformerEncoder.forward_infer for one stream."""

    def __init__(self, W: Dict[str, torch.Tensor], n_layers: int, n_head: int, d_model: int = 256):
        self.W, self.L, self.H, self.D = W, n_layers, n_head, d_model
        self.pe = sinusoid_table(5000, d_model)
        self.reset()

    def reset(self):
        self.states = None

    def _pos(self, x, offset):
        return x * math.sqrt(self.D) + self.pe[offset: offset + x.size(1)].unsqueeze(0)

    def _layer(self, l, x, mask, past_ctx, next_ctx, is_short):
        """contextual_block_encoder_layer.py:178-271; x (1, nb, S, D)."""
        W, p = self.W, f"encoder.encoders.{l}"
        nb = x.size(1)
        h = x.view(-1, x.size(-2), x.size(-1))
        m = mask.view(-1, mask.size(-2), mask.size(-1)) if mask is not None else None
        n1 = layer_norm(h, W[p + ".norm1.weight"], W[p + ".norm1.bias"])
        h = h + mha(W, p + ".self_attn", self.H, n1, n1, n1, m)
        n2 = layer_norm(h, W[p + ".norm2.weight"], W[p + ".norm2.bias"])
        h = h + ffn(W, p + ".feed_forward", n2)
        x = h.view(1, nb, h.size(-2), h.size(-1))
        if not is_short:
            if past_ctx is None:
                x[:, 0, 0, :] = x[:, 0, -1, :]
            else:
                x[:, 0, 0, :] = past_ctx[:, l, :]
            if nb > 1:
                x[:, 1:, 0, :] = x[:, 0:-1, -1, :]
            next_ctx[:, l, :] = x[:, -1, -1, :]
        return x

    def __call__(self, feats: torch.Tensor, is_final: bool) -> torch.Tensor:
        """feats (1, T, 80) -> encoder output (1, T_out, D) (possibly T_out == 0).
        contextual_block_transformer_encoder.py:241-419."""
        st = self.states
        if st is None:
            prev_addin = buf_before = buf_after = past_ctx = None
            n_proc = 0
        else:
            prev_addin, buf_before, buf_after = st["prev_addin"], st["buf_before"], st["buf_after"]
            n_proc, past_ctx = st["n_proc"], st["past_ctx"]
        xs = feats
        if st is not None:
            xs = torch.cat([buf_before, xs], dim=1)
        if is_final:
            buf_before = None
        else:
            n_samples = xs.size(1) // 4 - 1
            if n_samples < 2:
                self.states = dict(prev_addin=prev_addin, buf_before=xs, buf_after=buf_after,
                                   n_proc=n_proc, past_ctx=past_ctx)
                return xs.new_zeros(1, 0, self.D)
            n_res = xs.size(1) % 4 + 8
            buf_before = xs[:, xs.size(1) - n_res:]
            xs = xs[:, : n_samples * 4]
        xs = conv2d_subsample(self.W, xs)
        if buf_after is not None:
            xs = torch.cat([buf_after, xs], dim=1)
        total = xs.size(1)
        if is_final:
            block_num = math.ceil(float(total - (BLOCK - HOP_B - LOOK) - LOOK) / float(HOP_B))
            buf_after = None
        else:
            if total <= BLOCK:
                self.states = dict(prev_addin=prev_addin, buf_before=buf_before, buf_after=xs,
                                   n_proc=n_proc, past_ctx=past_ctx)
                return xs.new_zeros(1, 0, self.D)
            overlap = BLOCK - HOP_B
            block_num = max(0, total - overlap) // HOP_B
            res = total - HOP_B * block_num
            buf_after = xs[:, total - res:]
            xs = xs[:, : block_num * HOP_B + overlap]
        if n_proc == 0 and total <= BLOCK and is_final:       # short utterance, :345-351
            x = self._pos(xs, 0).unsqueeze(1)
            for l in range(self.L):
                x = self._layer(l, x, None, None, None, True)
            x = x.squeeze(0)
            self.states = None
            return layer_norm(x, self.W["encoder.after_norm.weight"], self.W["encoder.after_norm.bias"])
        block_num = max(block_num, 0)
        chunk = xs.new_zeros(1, block_num, BLOCK + 2, self.D)
        for i in range(block_num):
            cur = i * HOP_B
            clen = min(BLOCK, total - cur)
            data = xs[:, cur: cur + clen]
            addin = self._pos(data.mean(1, keepdim=True), i + n_proc)
            if prev_addin is None:
                prev_addin = addin
            chunk[:, i, 0] = prev_addin
            chunk[:, i, -1] = addin
            chunk[:, i, 1: clen + 1] = self._pos(data, cur + HOP_B * n_proc)
            prev_addin = addin
        mask = xs.new_zeros(1, block_num, BLOCK + 2, BLOCK + 2)
        mask[:, :, 1: BLOCK + 2, 0: BLOCK + 1] = 1
        next_ctx = xs.new_zeros(1, self.L, self.D)
        x = chunk
        for l in range(self.L):
            x = self._layer(l, x, mask, past_ctx, next_ctx, False)
        ys_chunk = x[:, :, 1: BLOCK + 1]
        offset = BLOCK - LOOK - HOP_B
        if is_final:
            y_len = xs.size(1) if n_proc == 0 else xs.size(1) - offset
        else:
            y_len = block_num * HOP_B + (offset if n_proc == 0 else 0)
        ys = ys_chunk.new_zeros(1, y_len, self.D)
        if n_proc == 0:
            ys[:, 0:offset] = ys_chunk[:, 0, 0:offset]
        for i in range(block_num):
            cur = i * HOP_B + (offset if n_proc == 0 else 0)
            if i == block_num - 1 and is_final:
                clen = min(BLOCK - offset, ys.size(1) - cur)
            else:
                clen = HOP_B
            ys[:, cur: cur + clen] = ys_chunk[:, i, offset: offset + clen]
        ys = layer_norm(ys, self.W["encoder.after_norm.weight"], self.W["encoder.after_norm.bias"])
        if is_final:
            self.states = None
        else:
            self.states = dict(prev_addin=prev_addin, buf_before=buf_before, buf_after=buf_after,
                               n_proc=n_proc + block_num, past_ctx=next_ctx)
        return ys
