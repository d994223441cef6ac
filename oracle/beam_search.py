"""Oracle: block-synchronous beam search with its hypothesis bookkeeping.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  speechcatcher/beam_search/beam_search.py:71-185   (two-pass scoring, pre-beam 40)
  speechcatcher/beam_search/beam_search.py:403-505  (extend_scorers, BBD repetition test)
  speechcatcher/beam_search/beam_search.py:507-838  (process_block, _decode_one_block, rewind)
  speechcatcher/beam_search/hypothesis.py:9-168
Quirks Q3-Q8 of SURVEY.md section 8 are reproduced literally.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import torch

from .ctc_prefix import CTCPrefixOracle
from .decoder import DecoderOracle
from .encoder import EncoderOracle

PRE_BEAM = 40
MAX_LENGTH = 500
BLOCK, HOP_B, LOOK = 40, 16, 16


@dataclass
class Hyp:
    yseq: List[int]
    score: float = 0.0
    scores: Dict[str, float] = field(default_factory=dict)
    states: Dict[str, Any] = field(default_factory=dict)
    xpos: List[int] = field(default_factory=list)


def _snapshot(hyps: List[Hyp]) -> List[Hyp]:
    """beam_search.py:358-401 -- states are never mutated in place by this port, so a
    shallow copy of the containers is an exact stand-in for the reference's deepcopy."""
    return [Hyp(list(h.yseq), h.score, dict(h.scores), dict(h.states), list(h.xpos)) for h in hyps]


class BeamSearchOracle:
    def __init__(self, encoder: EncoderOracle, decoder: DecoderOracle, ctc: CTCPrefixOracle,
                 beam_size: int, ctc_weight: float = 0.3, vocab: int = 1024, use_bbd: bool = False,
                 trace: Optional[Callable[[dict], None]] = None):
        self.encoder, self.decoder, self.ctc = encoder, decoder, ctc
        self.beam, self.V = beam_size, vocab
        self.w_dec, self.w_ctc = 1.0 - ctc_weight, ctc_weight
        self.sos = self.eos = vocab - 1
        self.use_bbd = use_bbd
        self.trace = trace
        self.reset()

    def reset(self):
        """beam_search.py:343-356, with the clean-reset semantics for the CTC store (Q2)."""
        self.enc_buf: Optional[torch.Tensor] = None
        self.running: Optional[List[Hyp]] = None
        self.prev_hyps: List[Hyp] = []
        self.processed_block = 0
        self.process_idx = 0
        self.encoder.reset()
        self.ctc.x, self.ctc.T = None, 0

    # -- scoring (beam_search.py:71-185) --------------------------------------
    def _score(self, hyps: List[Hyp], mem: torch.Tensor):
        n = len(hyps)
        ys = torch.tensor([h.yseq for h in hyps], dtype=torch.long)
        mem_b = mem.expand(n, -1, -1)
        dec, dec_states = self.decoder.batch_score(ys, [h.states.get("decoder") for h in hyps], mem_b)
        combined = torch.zeros(n, self.V)
        full = torch.zeros(n, self.V)
        combined += self.w_dec * dec
        full += self.w_dec * dec
        _, ids = torch.topk(full, k=min(PRE_BEAM, self.V), dim=-1)
        ctc, ctc_state = self.ctc.score_partial(ys, ids, [h.states.get("ctc") for h in hyps])
        combined += self.w_ctc * ctc
        return combined, dec, ctc, dec_states, ctc_state, ids

    def _has_repetition(self, hyps: List[Hyp]) -> bool:
        """beam_search.py:466-505."""
        for h in hyps:
            if len(h.yseq) < 2:
                continue
            last = h.yseq[-1]
            if last == self.sos or last == self.eos:
                continue
            if last in h.yseq[1:-1]:
                return True
        return False

    # -- one block (beam_search.py:655-838) -------------------------------------
    def _decode_one_block(self, mem: torch.Tensor, hyps: List[Hyp], is_final: bool) -> List[Hyp]:
        self.ctc.extend_prob(mem)
        ext = []
        for h in hyps:
            st = dict(h.states)
            if "ctc" in st:
                st["ctc"] = self.ctc.extend_state(st["ctc"])
            ext.append(Hyp(h.yseq, h.score, dict(h.scores), st, h.xpos))
        cur = ext
        if mem.size(1) > 0:
            prev_step = cur
            while self.process_idx < MAX_LENGTH:
                combined, dec, ctc, dec_states, ctc_state, ids = self._score(cur, mem)
                new = []
                for i, h in enumerate(cur):
                    top_s, top_t = torch.topk(combined[i], self.beam)
                    for s, tok in zip(top_s.tolist(), top_t.tolist()):
                        sc = dict(h.scores)
                        sc["decoder"] = sc.get("decoder", 0.0) + dec[i, tok].item()
                        sc["ctc"] = sc.get("ctc", 0.0) + ctc[i, tok].item()
                        new.append(Hyp(h.yseq + [tok], h.score + s, sc,
                                       {"decoder": dec_states[i],
                                        "ctc": self.ctc.select_state(ctc_state, i, tok)},
                                       h.xpos + [mem.size(1) - 1]))
                cur = sorted(new, key=lambda h: h.score, reverse=True)[: self.beam]
                if self.trace is not None:
                    self.trace(dict(process_idx=self.process_idx, T=mem.size(1), is_final=is_final,
                                    dec=dec, ctc=ctc, combined=combined, ids=ids,
                                    beam=[(list(h.yseq), h.score) for h in cur],
                                    cand=sorted((h.score for h in new), reverse=True)))
                if any(h.yseq[-1] == self.eos for h in cur):
                    if not is_final:
                        break
                    if max(cur, key=lambda h: h.score).yseq[-1] == self.eos:
                        break
                if self.use_bbd and not is_final and self._has_repetition(cur):
                    if len(prev_step) > 0:
                        cur = prev_step
                    break
                prev_step = cur
                if is_final and all(h.yseq[-1] == self.eos for h in cur):
                    break
                self.prev_hyps = _snapshot(cur)
                self.process_idx += 1
        if self.process_idx > 1 and len(self.prev_hyps) > 0:   # rewind, :827-836 (Q5)
            cur = self.prev_hyps
            self.process_idx -= 1
            self.prev_hyps = []
        return cur

    # -- one call (beam_search.py:507-653) ------------------------------------------
    def process_block(self, feats: torch.Tensor, is_final: bool) -> List[Hyp]:
        if self.running is None:
            self.running = [Hyp([self.sos], 0.0, {}, {}, [0])]
        if feats.size(1) < 3:                                   # :551-559 (Q12)
            enc_out = feats.new_zeros(1, 0, 256)
        else:
            enc_out = self.encoder(feats, is_final)
        self.last_enc_out = enc_out
        if enc_out.size(1) > 0:
            self.enc_buf = enc_out if self.enc_buf is None else torch.cat([self.enc_buf, enc_out], dim=1)
        cur, ret = self.running, None
        while True:
            end = BLOCK - LOOK + HOP_B * self.processed_block
            if self.enc_buf is not None and end < self.enc_buf.shape[1]:
                cur = ret = self._decode_one_block(self.enc_buf[:, :end], cur, False)
                self.processed_block += 1
            elif is_final and self.enc_buf is not None and self.enc_buf.shape[1] > 0:
                cur = ret = self._decode_one_block(self.enc_buf, cur, True)
                break
            else:
                break
        if ret is not None:
            self.running = ret
        return self.running
