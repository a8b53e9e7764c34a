"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/lwfb90.h declares; calls that need a device fail with an error code, not a crash."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

from lwfbrook90.jl_b200 import capi


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lwfb90.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lwfb_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    decl = declared_symbols()
    assert len(decl) >= 17
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/lwfb90.h but not exported"
    assert sorted(capi.SYMBOLS) == decl


def test_default_opts_mirror_the_reference_solver_settings():
    # src/LWFBrook90.jl:552-567: Tsit5, reltol 1e-5, (default) abstol 1e-6, dt 1e-3, dtmax 60/60/24
    o = capi.LwfbOpts()
    capi.lib().lwfb_default_opts(C.byref(o))
    assert o.reltol == 1e-5 and o.abstol == 1e-6 and o.dt_init == 1e-3 and o.dtmax == 60 / 60 / 24
    assert o.simulate_isotopes == 1 and o.record_daily == 1 and o.maxiters == 1000000


def test_create_rejects_bad_arguments_without_touching_a_device():
    L = capi.lib()
    h = C.c_void_p()
    o = capi.LwfbOpts()
    L.lwfb_default_opts(C.byref(o))
    o.n_layers, o.n_days = 0, 10
    assert L.lwfb_create(C.byref(h), C.byref(o)) == -4          # LWFB_ERR_UNSUPPORTED
    o.n_layers, o.n_days = 65, 10
    assert L.lwfb_create(C.byref(h), C.byref(o)) == -4
    o.n_layers, o.n_days = 5, 0
    assert L.lwfb_create(C.byref(h), C.byref(o)) == -1          # LWFB_ERR_ARG
    assert L.lwfb_create(None, C.byref(o)) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.LwfbError):
        capi.Handle(5, True, 0, 10)
