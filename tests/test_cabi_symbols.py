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


def _header_enum(prefix):
    """Names of one enum block of include/hbk.h, in declaration order (HBK_N_* terminator excluded)."""
    text = (ROOT / "include" / "hbk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(" + prefix + r"[A-Z0-9_]+)\b\s*(?:=\s*\d+)?\s*[,}]", text)
    seen = []
    for n in names:
        if n not in seen and not n.startswith("HBK_N_"):
            seen.append(n)
    return seen


def test_ctypes_binding_matches_the_header():
    """ABI drift guard: the ctypes structure has the layout gcc gives `hbk_structure`, and the slot lists of
    hydrobricks_b200/engine.py follow the header's enums entry by entry."""
    import subprocess
    import tempfile

    from hydrobricks_b200 import engine

    fields = [f for f, _ in engine.HbkStructure._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"hbk.h\"\nint main(void){\n"
    prog += 'printf("%zu\\n", sizeof(hbk_structure));\n'
    for f in fields:
        prog += f'printf("%zu\\n", offsetof(hbk_structure, {f}));\n'
    prog += 'printf("%d %d %d\\n", HBK_N_PARAMS, HBK_N_RECORDS, HBK_N_STATES);\nreturn 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = Path(tmp) / "layout.c", Path(tmp) / "layout"
        src.write_text(prog)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)])
        out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == ctypes.sizeof(engine.HbkStructure)
    for f, off in zip(fields, out[1:1 + len(fields)]):
        assert getattr(engine.HbkStructure, f).offset == int(off), f
    n_params, n_records, n_states = map(int, out[1 + len(fields):])
    assert (n_params, n_records, n_states) == (engine.N_PARAMS, len(engine.RECORD_SLOTS), len(engine.STATE_SLOTS))
    assert len(_header_enum("HBK_P_")) == n_params and len(_header_enum("HBK_R_")) == n_records
    assert len(_header_enum("HBK_S_")) == n_states
    # spot checks of the order (names differ in style, positions must agree)
    p, r, s = _header_enum("HBK_P_"), _header_enum("HBK_R_"), _header_enum("HBK_S_")
    assert p.index("HBK_P_CAPACITY") == engine.PARAM_SLOTS.index("capacity")
    assert p.index("HBK_P_K_ICE") == engine.PARAM_SLOTS.index("k_ice")
    assert p.index("HBK_P_GLACIER_SUBLIMATION") == engine.PARAM_SLOTS.index("glacier_sublimation")
    assert r.index("HBK_R_SLOW_WATER") == engine.RECORD_SLOTS.index("slow_water")
    assert r.index("HBK_R_GLACIER_SNOWPACK_SUBLIMATION") == engine.RECORD_SLOTS.index("glacier_snowpack_sublimation")
    assert s.index("HBK_S_QUICK") == engine.STATE_SLOTS.index("quick")
    assert s.index("HBK_S_AMT_SNOW_ICE") == engine.STATE_SLOTS.index("amt_snow_ice")
    # solver ids
    for name, const in (("rk4", "HBK_SOLVER_RK4"), ("crank_nicolson", "HBK_SOLVER_CRANK_NICOLSON"),
                        ("analytic_linear", "HBK_SOLVER_ANALYTIC_LINEAR")):
        m = re.search(const + r"\s*=\s*(\d+)", (ROOT / "include" / "hbk.h").read_text())
        assert m and int(m.group(1)) == engine.SOLVER[name]
