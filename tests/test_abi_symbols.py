"""The C-ABI library loads and exports every symbol include/valida_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "valida_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "valida_b200", "libvalida_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device(built):
    """Without a CUDA device the product path must fail loudly, never fall back to a CPU path."""
    import torch
    import pytest
    import valida_b200 as vb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VgpuError) as e:
        vb.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_references_oracle():
    """Nothing under valida_b200/ (the product) may import, link or name the oracle."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "valida_b200")):
        if "build" in d.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".inc")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"liboracle|oracle/|oracle_binding|import oracle|orc_", txt) and f != "build.py":
                    bad.append(os.path.join(d, f))
    assert not bad, bad
