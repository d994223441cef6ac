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
