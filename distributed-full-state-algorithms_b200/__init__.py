"""dfsa_b200: B200-native (sm_100a CUDA + NCCL/NVLink) implementation of the distributed full-state
algorithms behind the API of TysonRayJones/Distributed-Full-State-Algorithms.

The product is native: `libdfsa_b200.so` (CUDA kernels + transports, C-ABI in include/dfsa_b200.h) and the
drop-in C++ headers in `host/`. This Python package only builds those libraries in-tree and binds the host
API through ctypes (`api.py`) for tests and bench.py. There is no CPU fallback: without the built CUDA
library or without a GPU every call raises.

The directory name contains '-', so import it with importlib:
    dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
"""
from . import build as _build  # noqa: F401
from .api import (  # noqa: F401
    DeviceState,
    DfsaError,
    comm_end,
    comm_init,
    comm_rank,
    comm_size,
    comm_synch,
    device_lib,
    gate_fusion_enabled,
    host_lib,
    set_gate_fusion,
)

build = _build.build
