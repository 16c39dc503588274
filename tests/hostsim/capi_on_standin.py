"""TEST INFRASTRUCTURE (CPU): drives the extern "C" face of the host layer (host/dfsa_host_capi.cpp -- what api.py and bench.py call
through ctypes) on the CPU stand-in of the C-ABI, at DFSA_NP ranks (the stand-in forks them inside dfsa_host_comm_init, so every
rank runs this script). Checks the per-state entry points that need a live StateVector and therefore cannot be reached by the other
CPU tests: pendingGates / planPendingFlush (decoded with api.py's own decoder, compared with the stateless planner where the inputs
coincide), layout, flushGates, restoreLayout, and the amplitudes against a numpy dense simulation.

    DFSA_NP=4 python tests/hostsim/capi_on_standin.py        -> prints "capi on stand-in: ok ..." on rank 0, exit code 0
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
api = importlib.import_module("distributed-full-state-algorithms_b200.api")     # only its pure-Python helpers are used here


def main():
    lib = C.CDLL(os.path.join(HERE, "_build", "libdfsa_host_on_standin.so"))
    lib.dfsa_host_StateVector_new.restype = C.c_void_p
    lib.dfsa_host_state_planPendingFlush.restype = C.c_uint
    lib.dfsa_host_comm_getRank.restype = C.c_uint
    lib.dfsa_host_comm_getNumNodes.restype = C.c_uint
    lib.dfsa_host_plan_flush.restype = C.c_uint
    lib.dfsa_host_comm_init()
    rank, P = lib.dfsa_host_comm_getRank(), lib.dfsa_host_comm_getNumNodes()
    k = P.bit_length() - 1
    n = 8 + k
    L = n - k
    rng = np.random.default_rng(3)
    psi = C.c_void_p(lib.dfsa_host_StateVector_new(n))
    truth = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    lib.dfsa_host_state_setAllVecAmps(psi, np.ascontiguousarray(truth).ctypes.data_as(C.POINTER(C.c_double)))

    def layout():
        out = (C.c_uint * n)()
        lib.dfsa_host_state_layout(psi, out)
        return list(out)

    def apply_truth(t, ctrls, g):
        idx = np.arange(1 << n)
        sel = ((idx >> t) & 1) == 0
        for c in ctrls:
            sel &= ((idx >> c) & 1) == 1
        i0 = idx[sel]
        i1 = i0 | (1 << t)
        a0, a1 = truth[i0].copy(), truth[i1].copy()
        truth[i0] = g[0, 0] * a0 + g[0, 1] * a1
        truth[i1] = g[1, 0] * a0 + g[1, 1] * a1

    clock, last_use = 0, [0] * n
    for layer in range(3):
        gates = []
        where_before = layout()
        for q in range(n):
            for ctrls in ([], [(q + 2) % n]):
                g = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))[0]
                flat = np.ascontiguousarray(g).view(np.float64).reshape(-1)
                ptr = flat.ctypes.data_as(C.POINTER(C.c_double))
                if ctrls:
                    lib.dfsa_host_sv_manyCtrlOneTargGate(psi, (C.c_uint * 1)(*ctrls), 1, q, ptr)
                else:
                    lib.dfsa_host_sv_oneTargGate(psi, q, ptr)
                apply_truth(q, ctrls, g)
                gates.append((q, ctrls))
                clock += 1
                last_use[q] = clock
        assert lib.dfsa_host_state_pendingGates(psi) == len(gates), "every gate of the layer is deferred, rank-bit qubits included"
        assert layout() == where_before, "nothing moves until the queue is flushed"
        bufs = api._flush_plan_buffers(len(gates), n)
        steps, after = api._decode_flush_plan(lib.dfsa_host_state_planPendingFlush(psi, *bufs), *bufs, n)
        # the stateless planner (what tests/test_flush_plan.py drives) sees the same inputs and must give the same plan
        arr = (api.Gate1 * len(gates))()
        for i, (t, ctrls) in enumerate(gates):
            arr[i].target = t
            arr[i].ctrlMask = sum(1 << c for c in ctrls)
        bufs2 = api._flush_plan_buffers(len(gates), n)
        cnt2 = lib.dfsa_host_plan_flush((C.c_uint * n)(*where_before), n, L, (C.c_ulonglong * n)(*last_use), arr, len(gates), *bufs2)
        assert (steps, after) == api._decode_flush_plan(cnt2, *bufs2, n)
        relocations = [body for kind, body in steps if kind == "relocate"]
        assert (len(relocations) >= 1) == (P > 1) and all(1 <= len(r) <= max(k, 1) for r in relocations)
        assert sum(len(body) for kind, body in steps if kind == "gates") == len(gates)
        lib.dfsa_host_state_flushGates(psi)
        assert lib.dfsa_host_state_pendingGates(psi) == 0 and layout() == after
    out = np.empty(1 << n, dtype=np.complex128)
    lib.dfsa_host_state_getAllVecAmps(psi, out.ctypes.data_as(C.POINTER(C.c_double)))          # restores the layout
    assert layout() == list(range(n))
    err = float(np.max(np.abs(out - truth)))
    assert err < 1e-10, err
    lib.dfsa_host_state_delete(psi)
    if rank == 0:
        print("capi on stand-in: ok P=%d n=%d max|delta|=%.2e" % (P, n, err))
        sys.stdout.flush()
    lib.dfsa_host_comm_end()


if __name__ == "__main__":
    main()
