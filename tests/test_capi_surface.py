"""CPU tier: the C-ABI shared library loads, exports every symbol include/mpc_b200.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from mpc_limx_control_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_capi.SYMBOLS) == names


def test_params_struct_matches_header():
    L = _capi.lib()
    p = _capi.default_params()
    assert p.Ts == 0.005 and p.mass == 9.585 and abs(p.f_max - 2 * 9.585 * 9.8) < 1e-12
    assert list(p.q) == [1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1]
    assert p.gait_mpc_step == 5 and p.gait_dt == C.c_float(0.001).value and p.max_newton == 12
    assert L.mpc_b200_version() >= 100
    assert L.mpc_b200_strerror(-2).decode().startswith("no usable CUDA device")


def test_no_cpu_fallback():
    L = _capi.lib()
    if L.mpc_b200_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    p = _capi.default_params()
    assert L.mpc_b200_create(C.byref(p), 10, 16, 0, C.byref(h)) == _capi.ENODEV
    assert not h.value


def test_bad_arguments_rejected():
    L = _capi.lib()
    h = C.c_void_p()
    p = _capi.default_params()
    assert L.mpc_b200_create(C.byref(p), 7, 16, 0, C.byref(h)) == _capi.EINVAL      # unsupported horizon
    assert L.mpc_b200_create(None, 10, 16, 0, C.byref(h)) == _capi.EINVAL
    assert L.mpc_b200_tron1_default_params(None) == _capi.EINVAL


def test_product_does_not_reference_oracle():
    """The product tree must never import, link or execute anything under oracle/ or tests/."""
    bad = []
    for base in ("mpc_limx_control_b200", "include", "tools"):      # tools/ are measurement scripts: no oracle there either
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_lib|mpc_oracle|libmpc_oracle|emul_tron1|libemul", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
