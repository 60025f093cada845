"""Run test files of the reference's own Python test-suite (/root/reference/python/tests, build container only) with
its `_hydrobricks` extension replaced by a backend of tools/refpy.py: "b200" (this repo's host module) or "ref" (the
compiled reference, oracle/_ref). Nothing is written into the reference tree.

    python tools/run_reference_tests.py b200 test_parameters.py test_binding.py test_hydro_units.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import pytest

    import refpy

    backend, files = sys.argv[1], sys.argv[2:]
    refpy.load_reference_python(backend)
    scratch = ROOT / "gpurun_out"
    scratch.mkdir(exist_ok=True)
    tests = Path("/root/reference/python/tests")
    return pytest.main(["-q", "--no-header", "-p", "no:cacheprovider", "--rootdir", str(scratch), "--tb=line",
                        *[str(tests / f) for f in files]])


if __name__ == "__main__":
    sys.exit(main())
