"""The C-ABI library must load without a GPU and export every symbol include/hbk.h declares; calls that need a
device must fail loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "hbk.h").read_text()
    return sorted(set(re.findall(r"\b(hbk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(ROOT / "hydrobricks_b200" / "libhbk.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), name


def test_engine_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        return
    from hydrobricks_b200 import engine

    st = engine.HbkStructure(solver=2, n_soil_stores=1, time_step_days=1.0)
    try:
        engine.Engine(st, 4, 10)
    except engine.HbkError as err:
        assert "no CPU fallback" in str(err) or err.code == -2
    else:
        raise AssertionError("engine creation must fail without a CUDA device")


def test_product_never_imports_the_oracle():
    for path in (ROOT / "hydrobricks_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and path.is_file():
            text = path.read_text()
            assert "hb_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
