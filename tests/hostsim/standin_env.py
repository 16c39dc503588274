"""TEST INFRASTRUCTURE (CPU): points api.py's cached library handles at the CPU stand-in of the C-ABI, for the scripts of this directory
(bench_dry_run.py, cost_model_check.py). The host-only planners (dfsa_plan_*) still come from the real libdfsa_b200.so, which loads
without a GPU. THE PRODUCT HAS NO SUCH SWITCH: api.device_lib() / api.host_lib() only ever load the CUDA library; this module reaches
into their cache from the outside, in a test process."""
import ctypes as C
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
PKG = "distributed-full-state-algorithms_b200"


class DeviceLibOnStandIn:
    """attribute access like a ctypes.CDLL: host-only planners from the real library, everything else from the stand-in"""

    def __init__(self, standin, real):
        self._standin, self._real = standin, real

    def __getattr__(self, name):
        return getattr(self._real if name.startswith("dfsa_plan_") else self._standin, name)


def install():
    """-> (api module, stand-in CDLL); afterwards api.DeviceState & co. run on the stand-in in THIS process"""
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    api = importlib.import_module(PKG + ".api")
    standin = C.CDLL(os.path.join(HERE, "_build", "libdfsa_host_on_standin.so"), mode=C.RTLD_GLOBAL)
    real = C.CDLL(os.path.join(ROOT, PKG, "libdfsa_b200.so"))
    # the restypes api.device_lib() / api.host_lib() set on the real libraries
    standin.dfsa_last_error.restype = C.c_char_p
    standin.dfsa_version.restype = C.c_char_p
    standin.dfsa_comm_transport.restype = C.c_char_p
    standin.dfsa_stream_compute.restype = C.c_void_p
    standin.dfsa_state_ptr.restype = C.c_void_p
    standin.dfsa_state_ptr.argtypes = [C.c_void_p, C.c_int]
    standin.dfsa_state_num_amps_per_node.restype = C.c_uint64
    standin.dfsa_state_num_amps_per_node.argtypes = [C.c_void_p]
    standin.dfsa_launch_count.restype = C.c_uint64
    for name in ("dfsa_host_StateVector_new", "dfsa_host_DensityMatrix_new", "dfsa_host_dm_partialTrace", "dfsa_host_state_handle"):
        getattr(standin, name).restype = C.c_void_p
    standin.dfsa_host_state_numAmpsPerNode.restype = C.c_uint64
    standin.dfsa_host_state_getNorm2.restype = C.c_double
    standin.dfsa_host_comm_getRank.restype = C.c_uint
    standin.dfsa_host_comm_getNumNodes.restype = C.c_uint
    api._dev = DeviceLibOnStandIn(standin, real)
    api._host = standin
    standin.hostsim_remote_amps.restype = C.c_ulonglong
    return api, standin
