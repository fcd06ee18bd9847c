"""CPU-only checks of the boundary: the library loads, exports every symbol include/xped_b200.h declares, and
refuses to run without a GPU (no CPU fallback). No compute calls."""
import os
import re

import pytest

import xped_b200 as xb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xped_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = xb.load()
    declared = _declared_symbols()
    assert len(declared) > 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/xped_b200.h but not exported"
    assert sorted(xb.SIGNATURES) == declared, "ctypes signature table out of sync with the header"
    assert b"sm_100a" in lib.xb_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(xb.XbError) as e:
        xb.Context([("SU2", 0, 0)])
    assert e.value.status == 2   # XB_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xped_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "xped_oracle" not in src.replace("oracle/xped_oracle/rng.py", "") and "import oracle" not in src, f
