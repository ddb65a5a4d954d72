"""CPU tests of the drop-in boundary: the library loads, exports every symbol include/shems_b200.h declares,
and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "shems_b200.h")).read()
    return sorted(set(re.findall(r"SHEMS_API\s+[\w\s\*]+?\b(\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(sb):
    syms = header_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(sb._lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(sb._lib.SIGNATURES) == syms  # the Python binding covers exactly the header
    out = subprocess.run(["nm", "-D", "--defined-only", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    ours = {e for e in exported if e.startswith(("shems_", "replay_", "ddpg_"))}
    assert ours == set(syms), ours ^ set(syms)  # nothing undeclared leaks out


def test_sm100a_cubin_only(sb):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_params_and_errors_without_compute(sb, O, P98):
    p = sb.params_for_charger(98)
    for f, _ in p._fields_:
        assert getattr(p, f) == getattr(P98, f), f  # product constants == oracle constants (shems_LU1.jl:40-59, 92-99)
    for cid in (1, 2, 3, 4, 5, 6, 7, 8, 9, 97):
        a, b = sb.params_for_charger(cid), O.params_for_charger(cid)
        assert (a.b_soc_max, a.ev_soc_max, a.b_rate_max) == (b.b_soc_max, b.ev_soc_max, b.b_rate_max)
    with pytest.raises(KeyError):
        sb.params_for_charger(10)
    dp = sb.default_ddpg_params()
    assert (dp.l1, dp.l2, dp.batch) == (250, 500, 120) and dp.adam_eps == 1e-8


def test_no_cpu_fallback(sb):
    L = sb._lib
    if L.lib().shems_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ser = np.zeros((8, 100), np.float32)
    h = ctypes.c_void_p()
    p = sb.params_for_charger(98)
    st = L.lib().shems_create(ctypes.byref(p), ser.ctypes.data_as(L.PF), 100, 10, 4, 0, ctypes.byref(h))
    assert st == L.ERR_CUDA and b"no CPU fallback" in L.lib().shems_last_error()
    assert L.lib().replay_create(10, 0, ctypes.byref(h)) == L.ERR_CUDA
    dp = sb.default_ddpg_params()
    assert L.lib().ddpg_create(ctypes.byref(dp), 0, ctypes.byref(h)) == L.ERR_CUDA


def test_invalid_arguments(sb):
    L = sb._lib
    h = ctypes.c_void_p()
    p = sb.params_for_charger(98)
    ser = np.zeros((8, 72), np.float32)
    # nrows - maxsteps < 1: rand(1:(nrow-maxsteps)) is empty in the reference (shems_LU1.jl:225)
    assert L.lib().shems_create(ctypes.byref(p), ser.ctypes.data_as(L.PF), 72, 72, 1, 0, ctypes.byref(h)) == L.ERR_INVALID
    assert L.lib().shems_create(None, None, 72, 10, 1, 0, ctypes.byref(h)) == L.ERR_INVALID
    assert L.lib().shems_step(None, None, 0, None, None, None, None) == L.ERR_INVALID
    assert L.lib().replay_create(0, 0, ctypes.byref(h)) == L.ERR_INVALID


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "master-thesis-deep-reinforcement-learning-ddpg-in-home-energy-management_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt.replace("oracle/shems_oracle.c (oracle_philox)", "") and "import oracle" not in txt and "from oracle" not in txt, f
