"""CPU test: the C-ABI library loads without a GPU and exports every function include/dfsa_b200.h declares;
compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import product

ROOT = product.ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dfsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfsa_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = product.pkg().device_lib()
    names = declared_functions()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/dfsa_b200.h but not exported: %s" % missing


def test_host_library_exports_reference_api():
    h = product.pkg().host_lib()
    api = ["sv_oneTargGate", "sv_manyCtrlOneTargGate", "sv_swapGate", "sv_manyTargGate", "sv_pauliTensor", "sv_pauliGadget",
           "sv_phaseGadget", "dm_manyTargGate", "dm_swapGate", "dm_pauliTensor", "dm_pauliGadget", "dm_phaseGadget", "dm_krausMap",
           "dm_oneQubitDephasing", "dm_twoQubitDephasing", "dm_oneQubitDepolarising", "dm_twoQubitDepolarising", "dm_damping",
           "dm_expecPauliString", "dm_partialTrace", "comm_init", "comm_end", "comm_getRank", "comm_getNumNodes", "comm_synch"]
    for name in api:
        assert hasattr(h, "dfsa_host_" + name), name


def test_no_cpu_fallback_without_gpu():
    if product.gpu_count() > 0:
        pytest.skip("a GPU is present")
    dfsa = product.pkg()
    with pytest.raises(dfsa.DfsaError):
        dfsa.DeviceState("sv", 4)
    state = ctypes.c_void_p()
    rc = dfsa.device_lib().dfsa_state_create(0, 4, ctypes.byref(state))
    assert rc != 0 and b"no CPU fallback" in dfsa.device_lib().dfsa_last_error()
