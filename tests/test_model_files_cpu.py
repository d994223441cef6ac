"""CPU: model-directory handling and text assembly against goldens produced by the reference
(oracle/gen_golden_tokens.py)."""
import json
import shutil
import zipfile
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, model_dir
from speechcatcher_b200 import model_files as mf

G = json.loads((GOLDEN / "tokens.json").read_text())


def _dir_with_bpe(tmp_path, sub="m"):
    d = tmp_path / sub
    shutil.copytree(model_dir("m_d2"), d)
    shutil.copy(GOLDEN / "bpe_unigram1024.model", d / "bpe.model")
    return d


def test_token_list_matches_reference(tmp_path):
    tok, tl = mf.load_tokenizer(_dir_with_bpe(tmp_path))
    assert tl == G["token_list"] and tok.GetPieceSize() == 1024
    assert tl[0] == "<blank>" and tl[1] == "<unk>" and tl[-1] == "<sos/eos>" and "<s>" not in tl and "</s>" not in tl
    assert mf.load_tokenizer(model_dir("m_d2")) == (None, None)


def test_text_assembly_matches_reference():
    for call in G["calls"]:
        for text, toks, ids in call:
            assert mf.text_from_ids(ids, G["token_list"]) == (text, toks)
    assert mf.text_from_ids([5, 17], None) == ("5 17", ["5", "17"])


def test_oracle_replays_token_golden(tmp_path):
    from oracle.speech2text import OracleSpeech2Text
    from speechcatcher_b200.synthetic import make_model_dir, synth_audio
    d = make_model_dir(tmp_path / "m", "m_d2", seed=0, eos_bias=7.0)
    shutil.copy(GOLDEN / "bpe_unigram1024.model", Path(d) / "bpe.model")
    orc = OracleSpeech2Text(d, beam_size=5)
    audio = synth_audio(G["audio"]["seed"], G["audio"]["n"])
    for k, i in enumerate(range(0, len(audio), 8192)):
        fin = i + 8192 >= len(audio)
        got = orc(audio[i:i + 8192], is_final=fin, finalize_all=fin)
        assert [list(x) for x in got] == G["calls"][k]


def test_checkpoint_and_stats_search_order(tmp_path):
    d = tmp_path / "a" / "b" / "model"
    (d / "exp" / "asr_train").mkdir(parents=True)
    src = Path(model_dir("m_d2"))
    shutil.copy(src / "valid.acc.best.pth", d / "exp" / "asr_train" / "valid.acc.ave_6best.pth")
    assert mf.find_checkpoint(d) == d / "exp" / "asr_train" / "valid.acc.ave_6best.pth"
    shutil.copy(src / "valid.acc.best.pth", d / "model.pth")
    assert mf.find_checkpoint(d) == d / "model.pth"                     # the directory itself wins over exp/*
    assert mf.find_stats(d) == (None, None)
    alt = d.parent.parent / "asr_stats_raw_de_bpe1024" / "train"
    alt.mkdir(parents=True)
    mean, std = np.arange(80.0), np.full(80, 2.0)
    np.savez(alt / "feats_stats.npz", mean=mean, std=std)
    m, s = mf.find_stats(d)
    assert np.array_equal(m, mean) and np.array_equal(s, std) and m.dtype == np.float64
    shutil.copy(src / "feats_stats.npz", d / "feats_stats.npz")         # next to the checkpoint wins
    m2, s2 = mf.find_stats(d)
    st = np.load(src / "feats_stats.npz")
    np.testing.assert_array_equal(m2, st["sum"] / st["count"])
    with pytest.raises(FileNotFoundError):
        mf.find_checkpoint(tmp_path / "a")
    assert mf.state_dict_of({"state_dict": {"x": 1}}) == {"x": 1} and mf.state_dict_of({"x": 1}) == {"x": 1}


def test_resolve_model_dir(tmp_path):
    d = _dir_with_bpe(tmp_path)
    assert mf.resolve_model_dir(str(d)) == d
    z = tmp_path / "packed.zip"
    with zipfile.ZipFile(z, "w") as f:
        for p in d.iterdir():
            f.write(p, f"exp/asr_train_raw_de/{p.name}")
    got = mf.resolve_model_dir(str(z), cache_dir=str(tmp_path / "cache"))
    assert got.name == "asr_train_raw_de" and (got / "valid.acc.best.pth").exists() and (got / "config.yaml").exists()
    cache = tmp_path / "cache2"
    repo = mf.tags["de_streaming_transformer_xl"]
    shutil.copytree(d, cache / repo / "exp" / "asr")
    assert mf.find_checkpoint(mf.resolve_model_dir("de_streaming_transformer_xl", cache_dir=str(cache))).exists()
    with pytest.raises(FileNotFoundError, match="no downloader"):
        mf.resolve_model_dir("en_streaming_transformer_l", cache_dir=str(cache))
    with pytest.raises(ValueError):
        mf.load_model(str(d), decoder_impl="espnet")
