"""Oracle: streaming batched CTC prefix scorer (Watanabe et al., Algorithm 2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, for batch == 1 and
margin == 0 (the only configuration create_beam_search builds),
  speechcatcher/beam_search/ctc_prefix_score_full.py:35-86   (store)
  speechcatcher/beam_search/ctc_prefix_score_full.py:88-291  (__call__, partial scoring)
  speechcatcher/beam_search/ctc_prefix_score_full.py:293-368 (extend_prob / extend_state)
  speechcatcher/beam_search/scorers.py:117-146, 238-431      (CTCPrefixScorer wrapper)
including the quirk that rows appended after the first block are raw logits,
not log-softmax (scorers.py:349-350).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

LOGZERO = -10000000000.0


class CTCPrefixOracle:
    def __init__(self, W: Dict[str, torch.Tensor], blank: int = 0, eos: int = 1023):
        self.w, self.b = W["ctc.ctc_lo.weight"], W["ctc.ctc_lo.bias"]
        self.blank, self.eos = blank, eos
        self.x: Optional[torch.Tensor] = None      # (T, V): row t = emission scores of frame t
        self.T = 0

    # -- probability store -------------------------------------------------
    def extend_prob(self, enc: torch.Tensor) -> None:
        """enc (1, T, D): whole memory of the current block (scorers.py:330-350)."""
        logits = F.linear(enc, self.w, self.b)[0]
        if self.x is None:                                     # batch_init_state, :117-146
            self.x = torch.log_softmax(logits, dim=-1)
            self.T = self.x.size(0)
        elif self.T < logits.size(0):                          # ctc_prefix_score_full.py:302-324
            new = logits.clone()
            new[: self.T] = self.x
            self.x = new
            self.T = new.size(0)

    def extend_state(self, state):
        """state (r (T_old, 2), s) -> r grown to T rows with the blank-only path (:326-368)."""
        if self.x is None or state is None:
            return state
        r_prev, s_prev = state
        r_new = torch.full((self.T, 2), LOGZERO, dtype=torch.float32)
        start = min(max(r_prev.shape[0], 1), self.T)
        r_new[0:start] = r_prev[0:start]
        for t in range(start, self.T):
            r_new[t, 1] = r_new[t - 1, 1] + self.x[t, self.blank]
        return (r_new, s_prev)

    # -- scoring -------------------------------------------------------------
    def score_partial(self, yseqs: torch.Tensor, ids: torch.Tensor, states: List[Optional[Tuple]]):
        """yseqs (n, L) int64, ids (n, K) candidate tokens, states per hyp (r (T,2), s (V,)) or None.
        Returns scores (n, V) and the batched new state (r, log_psi, idmap)  (:88-291)."""
        n_bh, V, T = yseqs.size(0), self.x.size(1), self.T
        out_len = yseqs.size(1) - 1
        last = [int(v) for v in yseqs[:, -1]]
        K = ids.size(-1)
        if states[0] is None or states[0][0].shape[0] != T:    # scorers.py:284-299
            r_prev = torch.full((T, 2, n_bh), LOGZERO, dtype=torch.float32)
            r_prev[:, 1] = torch.cumsum(self.x[:, self.blank], 0).unsqueeze(1)
            s_prev = 0.0
        else:
            r_prev = torch.stack([s[0] for s in states], dim=2)
            s_prev = torch.stack([s[1] for s in states])
        idmap = torch.full((n_bh, V), -1, dtype=torch.long)
        idmap[torch.arange(n_bh).view(-1, 1), ids] = torch.arange(K)
        xn = self.x[:, ids.reshape(-1)].view(T, n_bh, K)        # x_[0]
        xb = self.x[:, self.blank].view(T, 1, 1).expand(T, n_bh, K)   # x_[1]
        x_ = torch.stack([xn, xb])                              # (2, T, n_bh, K)
        r = torch.full((T, 2, n_bh, K), LOGZERO, dtype=torch.float32)
        if out_len == 0:
            r[0, 0] = x_[0, 0]
        r_sum = torch.logsumexp(r_prev, 1)
        log_phi = r_sum.unsqueeze(2).repeat(1, 1, K)
        for i in range(n_bh):
            pos = idmap[i, last[i]]
            if pos >= 0:
                log_phi[:, i, pos] = r_prev[:, 1, i]
        start = min(max(out_len, 1), T)
        for t in range(start, T):
            rp = r[t - 1]
            rr = torch.stack([rp[0], log_phi[t - 1], rp[0], rp[1]]).view(2, 2, n_bh, K)
            r[t] = torch.logsumexp(rr, 1) + x_[:, t]
        log_phi_x = torch.cat((log_phi[0].unsqueeze(0), log_phi[:-1]), dim=0) + x_[0]
        log_psi = torch.full((n_bh, V), LOGZERO, dtype=torch.float32)
        log_psi_ = torch.logsumexp(torch.cat((log_phi_x[start:T], r[start - 1, 0].unsqueeze(0)), dim=0), dim=0)
        for i in range(n_bh):
            log_psi[i, ids[i]] = log_psi_[i]
        for i in range(n_bh):
            log_psi[i, self.eos] = r_sum[T - 1, i]
        log_psi[:, self.blank] = LOGZERO
        return (log_psi - s_prev), (r, log_psi, idmap)

    @staticmethod
    def select_state(state, i: int, tok: int):
        """scorers.py:382-431."""
        r, log_psi, idmap = state
        s = log_psi[i, tok].expand(log_psi.size(1))
        k = idmap[i, tok]
        return (r[:, :, i, k] if k >= 0 else r[:, :, i, 0], s)
