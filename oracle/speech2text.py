"""Oracle: the stream facade (one object per stream, like the reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  speechcatcher/speech2text_streaming.py:43-155 (construction), :252-263 (reset),
  :402-539 (__call__ incl. the output filter that hard-codes EOS id 1023).
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import yaml

from .beam_search import BeamSearchOracle, Hyp
from .ctc_prefix import CTCPrefixOracle
from .decoder import DecoderOracle
from .encoder import EncoderOracle
from .frontend import FrontendOracle, load_stats


def load_model_dir(model_dir):
    """(weights dict, config dict, mean, std) from an ESPnet-style model dir
    (speech2text_streaming.py:157-250, :76-95)."""
    model_dir = Path(model_dir)
    ckpt = torch.load(model_dir / "valid.acc.best.pth", map_location="cpu")
    W = {k: v.to(torch.float32) for k, v in ckpt.get("model", ckpt).items()}
    cfg = yaml.safe_load(open(model_dir / "config.yaml"))
    mean = std = None
    if (model_dir / "feats_stats.npz").exists():
        mean, std = load_stats(model_dir / "feats_stats.npz")
    return W, cfg, mean, std


def load_token_list(model_dir):
    """speech2text_streaming.py:97-124: ESPnet vocabulary from the SentencePiece model next to the checkpoint,
    ["<blank>", SP[0], SP[3..n-1], "<sos/eos>"]; None without a bpe.model (same three search locations)."""
    d = Path(model_dir)
    for p in (d / "bpe.model", d.parent.parent / "data/de_token_list/bpe_unigram1024/bpe.model",
              d / "../data/de_token_list/bpe_unigram1024/bpe.model"):
        if p.exists():
            import sentencepiece as spm
            sp = spm.SentencePieceProcessor()
            sp.Load(str(p))
            n = sp.GetPieceSize()
            return ["<blank>", sp.IdToPiece(0)] + [sp.IdToPiece(i) for i in range(3, n)] + ["<sos/eos>"]
    return None


class OracleSpeech2Text:
    def __init__(self, model_dir, beam_size: int = 5, ctc_weight: float = 0.3, use_bbd: bool = False,
                 trace: Optional[Callable[[dict], None]] = None):
        W, cfg, mean, std = load_model_dir(model_dir)
        enc, dec = cfg.get("encoder_conf", {}), cfg.get("decoder_conf", {})
        d_model = enc.get("output_size", 256)
        V = W["decoder.embed.0.weight"].shape[0]
        self.frontend = FrontendOracle(mean, std)
        self.encoder = EncoderOracle(W, enc.get("num_blocks", 12), enc.get("attention_heads", 4), d_model)
        self.decoder = DecoderOracle(W, dec.get("num_blocks", 6), dec.get("attention_heads", 4), d_model)
        self.ctc = CTCPrefixOracle(W, blank=0, eos=V - 1)
        self.search = BeamSearchOracle(self.encoder, self.decoder, self.ctc, beam_size, ctc_weight, V,
                                       use_bbd, trace)
        self.beam_size = beam_size
        self.last_feats: Optional[torch.Tensor] = None
        self.token_list = load_token_list(model_dir)
        self.reset()

    def reset(self):
        self.frontend.reset()
        self.search.reset()
        self.hyps: Optional[List[Hyp]] = None

    @torch.no_grad()
    def __call__(self, speech, is_final: bool = False, finalize_all: bool = False
                 ) -> List[Tuple[str, List[str], List[int]]]:
        if isinstance(speech, np.ndarray):
            speech = torch.from_numpy(speech)
        speech = speech.to(torch.float32)
        if speech.dim() == 1:
            feats = self.frontend(speech, is_final)
        elif speech.dim() == 2:                      # :438-446 pre-computed features: numpy normalisation, no frontend
            f = speech.numpy()
            if self.frontend.mean is not None:
                f = (f - self.frontend.mean) / self.frontend.std
            feats = torch.from_numpy(np.asarray(f)).to(torch.float32)
        else:                                        # :447-450 already batched (1, T, 80): used as is
            assert speech.size(0) == 1
            feats = speech[0]
        self.last_feats = feats
        if feats is None:
            return []
        self.hyps = self.search.process_block(feats.unsqueeze(0), is_final)
        if not is_final or not finalize_all:
            out = [h for h in self.hyps if h.yseq[-1] == 1023]
        else:
            out = self.hyps
        results = []
        for h in out:
            if is_final:
                ids = h.yseq[1:]
                if len(ids) > 0 and ids[-1] == 1023:
                    ids = ids[:-1]
            else:
                ids = h.yseq[1:1]              # output_index is always 0 here
            ids = [t for t in ids if t not in (0, 1, 1023)]
            if self.token_list is not None:                      # :520-523
                toks = [self.token_list[t] for t in ids]
                text = "".join(toks).replace("\u2581", " ").strip()
            else:                                                # :531-535
                toks = [str(t) for t in ids]
                text = " ".join(toks)
            results.append((text, toks, list(ids)))
        return results
