"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol include/topopt_b200.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "topopt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tob_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from topopt_in_petsc_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, set(_lib.SIGNATURES) ^ set(syms)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from topopt_in_petsc_b200 import Context, TobError
    with pytest.raises(TobError):
        Context(9, 5, 5)


def test_sass_is_blackwell_native():
    """The hot kernel must contain TMA loads (UTMALDG) and FP64 math, built for sm_100a."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "topopt_in_petsc_b200", "libtopopt_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    i = out.find("k_hex8_zmarch")
    assert i >= 0
    body = out[i:i + 400000]
    assert "UTMALDG" in body and "DFMA" in body and "SYNCS" in body
