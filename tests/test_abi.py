"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol that
include/sgp.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spark_gp_b200 import build, _native
    build.build_native()
    return _native.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgp_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from spark_gp_b200 import _native
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert sorted(_native.EXPORTS) == declared          # binding and header agree
    for name in declared:
        assert hasattr(lib, name), name


def test_version(lib):
    assert lib.sgp_version() >= 100


def test_struct_layout_matches_header():
    from spark_gp_b200 import _native
    assert C.sizeof(_native.KernelTerm) == 32            # int32,int32,double,double,pointer
    assert C.sizeof(_native.KernelDesc) == 16


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.sgp_ctx_create(C.byref(h), 0)
    assert rc == 2                                       # SGP_E_CUDA
    assert b"no CPU fallback" in lib.sgp_last_error(None)
    import spark_gp_b200 as sg
    with pytest.raises(sg.SgpError):
        sg.ProjectedProcessEngine(0)


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "spark_gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f == "__init__.py", f


def test_jni_glue_syntax_checks_against_stub_header():
    """integration/jni/sgp_jni.cpp is guarded by __has_include(<jni.h>) (no JDK in this image); with the stand-in header of
    tests/support/jni_stub it must at least parse and type-check against include/sgp.h."""
    import shutil, subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        import pytest
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(root, "tests", "support", "jni_stub"),
                        "-I" + os.path.join(root, "include"), os.path.join(root, "integration", "jni", "sgp_jni.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
