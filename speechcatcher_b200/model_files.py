"""Locating and reading the files of an ESPnet-style model directory, and turning token ids into text
(SURVEY.md section 8(f), row N4).  Pure host code, no CUDA.

Mirrors the reference's search orders and the token-list construction:
  checkpoint            speechcatcher/speech2text_streaming.py:163-189
  feats_stats.npz       speechcatcher/speech2text_streaming.py:76-95, model/checkpoint_loader.py:210-237
  bpe.model, token list speechcatcher/speech2text_streaming.py:100-124
  ids -> text           speechcatcher/speech2text_streaming.py:520-535
  load_model(tag, ...)  speechcatcher/speechcatcher.py:126-227 (native decoder branch)
The German-specific fall-back locations (`asr_stats_raw_de_bpe1024`, `data/de_token_list/bpe_unigram1024`) are the
reference's: English / Spanish checkpoints only get MVN statistics and a tokenizer when the files sit next to the
checkpoint (SURVEY.md N4 caveat) -- kept as is.
"""
from __future__ import annotations

import hashlib
import os
import shutil
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

CHECKPOINT_NAMES = ["valid.acc.best.pth", "valid.acc.ave_6best.pth", "valid.acc.ave.pth", "model.pth", "checkpoint.pth"]

# short tag -> Hugging Face repo id (speechcatcher.py:50-57)
tags = {
    "de_streaming_transformer_m": "speechcatcher/speechcatcher_german_espnet_streaming_transformer_13k_train_size_m_raw_de_bpe1024",
    "de_streaming_transformer_l": "speechcatcher/speechcatcher_german_espnet_streaming_transformer_13k_train_size_l_raw_de_bpe1024",
    "de_streaming_transformer_xl": "speechcatcher/speechcatcher_german_espnet_streaming_transformer_26k_train_size_xl_raw_de_bpe1024",
    "es_streaming_transformer_m": "speechcatcher/wordcab_speechcatcher_spanish_espnet_streaming_transformer_35k_train_size_m_raw_es_bpe1024",
    "es_streaming_transformer_l": "speechcatcher/wordcab_speechcatcher_spanish_espnet_streaming_transformer_35k_train_size_l_raw_es_bpe1024",
    "en_streaming_transformer_m": "speechcatcher/wordcab_speechcatcher_english_espnet_streaming_transformer_35k_train_size_m_raw_en_bpe1024",
    "en_streaming_transformer_l": "speechcatcher/wordcab_speechcatcher_english_espnet_streaming_transformer_35k_train_size_l_raw_en_bpe1024",
}


def find_checkpoint(model_dir) -> Path:
    model_dir = Path(model_dir)
    paths = [model_dir / n for n in CHECKPOINT_NAMES]
    for exp in model_dir.glob("exp/*/"):
        paths += [exp / n for n in CHECKPOINT_NAMES]
    for p in paths:
        if p.exists():
            return p
    raise FileNotFoundError(f"No checkpoint found in {model_dir}")


def state_dict_of(checkpoint: Dict) -> Dict:
    """`{"model": sd}`, `{"state_dict": sd}` or a bare state dict (checkpoint_loader.py:172-178)."""
    if "model" in checkpoint:
        return checkpoint["model"]
    if "state_dict" in checkpoint:
        return checkpoint["state_dict"]
    return checkpoint


def stats_search_paths(model_dir) -> List[Path]:
    d = Path(model_dir)
    return [d / "feats_stats.npz", d.parent / "asr_stats_raw_de_bpe1024/train/feats_stats.npz",
            d.parent.parent / "asr_stats_raw_de_bpe1024/train/feats_stats.npz", d / "../stats/train/feats_stats.npz"]


def bpe_search_paths(model_dir) -> List[Path]:
    d = Path(model_dir)
    return [d / "bpe.model", d.parent.parent / "data/de_token_list/bpe_unigram1024/bpe.model",
            d / "../data/de_token_list/bpe_unigram1024/bpe.model"]


def read_stats(path) -> Tuple[np.ndarray, np.ndarray]:
    st = np.load(path)
    if "mean" in st:
        mean, std = st["mean"], st["std"]
    elif "sum" in st and "sum_square" in st and "count" in st:
        count = st["count"]
        mean = st["sum"] / count
        std = np.sqrt(np.maximum(st["sum_square"] / count - mean ** 2, 1e-10))
    else:
        raise ValueError(f"Unknown stats format. Keys: {list(st.keys())}")
    return np.ascontiguousarray(mean, np.float64), np.ascontiguousarray(std, np.float64)


def find_stats(model_dir) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """First readable feats_stats.npz in the reference's search order, else (None, None): features stay unnormalised."""
    for p in stats_search_paths(model_dir):
        if p.exists():
            try:
                return read_stats(p)
            except Exception:
                continue
    return None, None


def load_tokenizer(model_dir):
    """(SentencePieceProcessor | None, token_list | None).  ESPnet's vocabulary drops SentencePiece's <s> (1) and
    </s> (2): ["<blank>", SP[0], SP[3..n-1], "<sos/eos>"]."""
    try:
        import sentencepiece as spm
    except ImportError:
        return None, None
    for p in bpe_search_paths(model_dir):
        if p.exists():
            try:
                tok = spm.SentencePieceProcessor()
                tok.Load(str(p))
                n = tok.GetPieceSize()
                return tok, (["<blank>", tok.IdToPiece(0)] + [tok.IdToPiece(i) for i in range(3, n)] + ["<sos/eos>"])
            except Exception:
                continue
    return None, None


def text_from_ids(token_ids, token_list=None) -> Tuple[str, List[str]]:
    """(text, tokens) of already filtered token ids; without a token list the ids themselves are the tokens."""
    if token_list is not None:
        toks = [token_list[int(t)] for t in token_ids]
        return "".join(toks).replace("▁", " ").strip(), toks
    toks = [str(int(t)) for t in token_ids]
    return " ".join(toks), toks


def resolve_model_dir(tag: str, cache_dir: str = "~/.cache/espnet") -> Path:
    """Where the model of `tag` lives on this machine.  `tag`: a model directory, a packed archive (zip / tar), a short
    tag from `tags`, or a Hugging Face repo id.  There is no downloader here: repo ids are looked up in `cache_dir`
    (where espnet_model_zoo / a previous reference run unpacked them); a miss raises with the paths that were tried."""
    p = Path(os.path.expanduser(str(tag)))
    if p.is_dir():
        return _dir_with_checkpoint(p)
    cache = Path(os.path.expanduser(cache_dir))
    if p.is_file():
        out = cache / ("unpacked_" + hashlib.sha1(str(p.resolve()).encode()).hexdigest()[:16])
        if not out.exists():
            out.mkdir(parents=True)
            shutil.unpack_archive(str(p), str(out))
        return _dir_with_checkpoint(out)
    repo = tags.get(tag, tag)
    tried = []
    for cand in (cache / repo, cache / repo.replace("/", "--"), cache / ("models--" + repo.replace("/", "--")),
                 cache / repo.split("/")[-1]):
        tried.append(str(cand))
        if cand.is_dir():
            try:
                return _dir_with_checkpoint(cand)
            except FileNotFoundError:
                continue
    raise FileNotFoundError(f"model '{tag}' is not on this machine (looked in {tried}); speechcatcher_b200 has no "
                            f"downloader -- unpack the ESPnet model archive and pass its directory")


def _dir_with_checkpoint(root: Path) -> Path:
    """The directory below `root` that holds the checkpoint (the reference takes the parent of `asr_model_file`)."""
    try:
        return find_checkpoint(root).parent      # config.yaml, feats_stats.npz and bpe.model are looked up beside it
    except FileNotFoundError:
        pass
    for name in CHECKPOINT_NAMES:
        hits = sorted(root.rglob(name))
        if hits:
            return hits[0].parent
    raise FileNotFoundError(f"No checkpoint found below {root}")


def load_model(tag, device="cuda", beam_size=5, quiet=False, cache_dir="~/.cache/espnet", decoder_impl="native",
               fp16=False, use_bbd=False, **engine_kw):
    """speechcatcher.py:126 for the native decoder: resolve the model directory and construct the streaming facade.
    `fp16` is accepted and ignored like in the reference (:208-214 turns it off); the tensor-core mode of this path is
    `dtype="bfloat16"` (pass it through `engine_kw`)."""
    if decoder_impl != "native":
        raise ValueError("speechcatcher_b200 implements the native decoder only (decoder_impl='native')")
    from .speech2text_streaming import Speech2TextStreaming
    model_dir = resolve_model_dir(tag, cache_dir)
    if not quiet:
        print(f"Loading model from {model_dir}")
    return Speech2TextStreaming(model_dir=model_dir, beam_size=beam_size, ctc_weight=0.3, device=device,
                                use_bbd=use_bbd, **engine_kw)
