"""No-GPU checks of the C-ABI library: it loads, exports every symbol include/rapidnet_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from rapidnet_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rapidnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_match_binding_list():
    assert _declared_symbols() == sorted(cabi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} not exported by {cabi.LIB_PATH}"


def test_buffer_enum_matches_header():
    text = open(os.path.join(ROOT, "include", "rapidnet_b200.h")).read()
    body = text[text.index("typedef enum rn_buffer_id"):text.index("} rn_buffer_id;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for n in re.findall(r"RN_BUF_([A-Z_0-9]+)", body) if n != "COUNT_"]
    assert names == cabi.BUFFER_IDS


def test_no_cpu_fallback(toy):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path cannot be exercised")
    prob, _, _ = toy
    with pytest.raises(cabi.RapidNetError) as ei:
        cabi.Solver(prob)
    assert "RN_ERR_CUDA" in str(ei.value) and "no CPU path" in str(ei.value)


def test_null_arguments_rejected():
    lib = cabi.load()
    h = C.c_void_p()
    assert lib.rn_create(None, None, None, None, 0, C.byref(h)) == 1
    assert lib.rn_factor_step(None) == 1
    assert lib.rn_destroy(None) == 0
