"""CPU: the C-ABI shared library loads and exports every symbol include/speechcatcher_b200.h declares."""
import re
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def test_header_symbols_exported_and_bound():
    from speechcatcher_b200 import _lib
    lib = _lib.load()
    header = (REPO / "include" / "speechcatcher_b200.h").read_text()
    declared = set(re.findall(r"\b(sc_[A-Za-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert b"sm_100a" in lib.sc_version()


def test_product_path_does_not_import_oracle():
    """The shipped package must never route through the CPU oracle."""
    for py in (REPO / "speechcatcher_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py


def test_missing_cuda_fails_loudly():
    import pytest
    import torch
    from helpers import model_dir
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from speechcatcher_b200 import StreamGroup
    with pytest.raises(RuntimeError):
        StreamGroup(model_dir("m_d2"), n_streams=1)


def test_workspace_of_every_baseline_config_fits_one_b200():
    """Host-only arithmetic (no GPU call): the caller-owned workspace of each BASELINE.json configuration, per GPU,
    against 180 GB of HBM3e."""
    import ctypes as C
    from speechcatcher_b200 import _lib
    from speechcatcher_b200._lib import ScConfig
    lib = _lib.load()

    def gib(**kw):
        base = dict(d_model=256, enc_heads=8, enc_layers=30, dec_heads=8, dec_layers=14, vocab=1024, ffn=2048,
                    n_streams=256, beam=10, max_chunk=8192, max_frames=int(61 * 25) + 64, use_bbd=0, precision=1,
                    ctc_weight=0.3)
        base.update(kw)
        n = C.c_size_t()
        _lib.check(lib.sc_engine_workspace_bytes(C.byref(ScConfig(**base)), C.byref(n)))
        return n.value / 2 ** 30

    xl_bf16, xl_f32 = gib(), gib(precision=0)
    assert 20 < xl_bf16 < 40 and xl_bf16 < xl_f32 < 70              # config 2: bf16 KV caches halve the big buffers
    assert abs(4 * gib(n_streams=64) - xl_bf16) < 1.0               # 4 shards of 64 streams = one group of 256
    assert gib(enc_layers=18, dec_layers=8) < xl_bf16               # config 3 (L, 256 streams per GPU)
    assert gib(enc_layers=18, dec_layers=8, n_streams=60, max_frames=185 * 25 + 64) < 20   # config 4: 60 segments <= 180 s
    assert gib(beam=20) < 80                                         # config 5
    for bad in (dict(beam=21), dict(d_model=512), dict(n_streams=0), dict(max_frames=5000)):
        n = C.c_size_t()
        base = dict(d_model=256, enc_heads=8, enc_layers=30, dec_heads=8, dec_layers=14, vocab=1024, ffn=2048,
                    n_streams=256, beam=10, max_chunk=8192, max_frames=1589, use_bbd=0, precision=1, ctc_weight=0.3)
        base.update(bad)
        assert lib.sc_engine_workspace_bytes(C.byref(ScConfig(**base)), C.byref(n)) != 0
        assert lib.sc_last_error()
