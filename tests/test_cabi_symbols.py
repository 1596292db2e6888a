"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/cmoe_b200.h declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cmoe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmoe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cornell_moe_b200 import capi
    lib = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/cmoe_b200.h but not exported: {missing}"


def test_no_cpu_fallback():
    from cornell_moe_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(capi.NoDeviceError):
        capi.cholesky(np.eye(3))
    with pytest.raises(capi.NoDeviceError):
        capi.GaussianProcess(0, 1.0, [1.0], np.zeros((2, 1)) + [[0.0], [1.0]], [0.0, 1.0], [0.1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cornell-moe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "moe_oracle" not in txt and "import oracle" not in txt and "libmoe_ref" not in txt, f
