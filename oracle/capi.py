"""ctypes view of oracle/libdfsa_oracle.so (the plain-C restatement, dfsa_oracle.c). CHECKER ONLY.

`OracleState` exposes one method per reference API function, named exactly like the op names used
by tests (sv_oneTargGate, dm_damping, ...), so a test can replay the same op list on the oracle and
on the CUDA product and compare.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdfsa_oracle.so")
_lib = None


def build():
    """Compile the C restatement (and the real reference where /root/reference exists)."""
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "all"], check=True, env=env, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "dfsa_oracle.c")):
            env = dict(os.environ)
            env.pop("CC", None)
            subprocess.run(["make", "-C", _HERE, "libdfsa_oracle.so"], check=True, env=env, stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_dm_partialTrace.restype = C.c_void_p
        for name in ("orc_log_amps_per_node", "orc_num_qubits"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _ints(xs):
    xs = [int(x) for x in xs]
    return (C.c_int * max(len(xs), 1))(*xs), len(xs)


def _dbl(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class OracleState:
    """P virtual ranks of the reference's StateVector / DensityMatrix (src/states.hpp:13-69)."""

    def __init__(self, kind, num_qubits, num_nodes=1, _handle=None):
        self.kind = kind
        self.num_qubits = int(num_qubits)
        self.num_nodes = int(num_nodes)
        self.h = C.c_void_p(_handle) if _handle is not None else C.c_void_p(lib().orc_create(int(kind == "dm"), self.num_qubits, self.num_nodes))
        if not self.h:
            raise ValueError("orc_create failed (ranks must be a power of 2 and <= 2^qubits)")
        self.total_bits = (2 if kind == "dm" else 1) * self.num_qubits
        self.values = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    # ---- state I/O
    def set_amps(self, amps):
        a, p = _dbl(amps)
        assert a.size == 1 << self.total_bits
        lib().orc_set_amps(self.h, p)

    def get_amps(self):
        out = np.empty(1 << self.total_bits, dtype=np.complex128)
        lib().orc_get_amps(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def init_hash(self, seed):
        lib().orc_init_hash(self.h, C.c_uint64(seed))

    # ---- state-vector API
    def sv_oneTargGate(self, target, gate):
        g, p = _dbl(gate)
        lib().orc_sv_oneTargGate(self.h, int(target), p)

    def sv_manyCtrlOneTargGate(self, ctrls, target, gate):
        g, p = _dbl(gate)
        c, n = _ints(ctrls)
        lib().orc_sv_manyCtrlOneTargGate(self.h, c, n, int(target), p)

    def sv_swapGate(self, q1, q2):
        lib().orc_sv_swapGate(self.h, int(q1), int(q2))

    def sv_manyTargGate(self, targets, gate):
        g, p = _dbl(gate)
        t, n = _ints(targets)
        lib().orc_sv_manyTargGate(self.h, t, n, p)

    def sv_pauliTensor(self, targets, paulis):
        t, n = _ints(targets)
        q, _ = _ints(paulis)
        lib().orc_sv_pauliTensor(self.h, t, q, n)

    def sv_pauliGadget(self, targets, paulis, theta):
        t, n = _ints(targets)
        q, _ = _ints(paulis)
        lib().orc_sv_pauliGadget(self.h, t, q, n, C.c_double(theta))

    def sv_phaseGadget(self, targets, theta):
        t, n = _ints(targets)
        lib().orc_sv_phaseGadget(self.h, t, n, C.c_double(theta))

    # ---- density-matrix API
    def dm_manyTargGate(self, targets, gate):
        g, p = _dbl(gate)
        t, n = _ints(targets)
        lib().orc_dm_manyTargGate(self.h, t, n, p)

    def dm_swapGate(self, q1, q2):
        lib().orc_dm_swapGate(self.h, int(q1), int(q2))

    def dm_pauliTensor(self, targets, paulis):
        t, n = _ints(targets)
        q, _ = _ints(paulis)
        lib().orc_dm_pauliTensor(self.h, t, q, n)

    def dm_pauliGadget(self, targets, paulis, theta):
        t, n = _ints(targets)
        q, _ = _ints(paulis)
        lib().orc_dm_pauliGadget(self.h, t, q, n, C.c_double(theta))

    def dm_phaseGadget(self, targets, theta):
        t, n = _ints(targets)
        lib().orc_dm_phaseGadget(self.h, t, n, C.c_double(theta))

    def dm_krausMap(self, targets, kraus_ops):
        k, p = _dbl(np.stack([np.asarray(m, dtype=np.complex128) for m in kraus_ops]))
        t, n = _ints(targets)
        lib().orc_dm_krausMap(self.h, p, len(kraus_ops), t, n)

    def dm_oneQubitDephasing(self, q, prob):
        lib().orc_dm_oneQubitDephasing(self.h, int(q), C.c_double(prob))

    def dm_twoQubitDephasing(self, q1, q2, prob):
        lib().orc_dm_twoQubitDephasing(self.h, int(q1), int(q2), C.c_double(prob))

    def dm_oneQubitDepolarising(self, q, prob):
        lib().orc_dm_oneQubitDepolarising(self.h, int(q), C.c_double(prob))

    def dm_twoQubitDepolarising(self, q1, q2, prob):
        lib().orc_dm_twoQubitDepolarising(self.h, int(q1), int(q2), C.c_double(prob))

    def dm_damping(self, q, prob):
        lib().orc_dm_damping(self.h, int(q), C.c_double(prob))

    def dm_expecPauliString(self, coeffs, paulis):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        p, n = _ints(np.asarray(paulis).reshape(-1))
        assert n == coeffs.size * self.num_qubits
        out = (C.c_double * 2)()
        lib().orc_dm_expecPauliString(self.h, coeffs.ctypes.data_as(C.POINTER(C.c_double)), coeffs.size, p, out)
        v = complex(out[0], out[1])
        self.values.append(v)
        return v

    def dm_partialTrace(self, targets):
        """Returns the reduced state; like the reference, `self` is mutated on the relocation path."""
        t, n = _ints(targets)
        h = lib().orc_dm_partialTrace(self.h, t, n)
        if not h:
            raise ValueError("partialTrace precondition failed")
        return OracleState("dm", self.num_qubits - n, self.num_nodes, _handle=h)


def superoperator(kraus_ops):
    """src/misc.hpp:58-81"""
    k = np.stack([np.asarray(m, dtype=np.complex128) for m in kraus_ops])
    d = k.shape[1]
    nt = d.bit_length() - 1
    out = np.zeros((d * d, d * d), dtype=np.complex128)
    lib().orc_superoperator(k.ctypes.data_as(C.POINTER(C.c_double)), len(kraus_ops), nt, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def plan_manyTarg(log_amps_per_node, targets):
    t, n = _ints(targets)
    out = (C.c_int * n)()
    lib().orc_plan_manyTarg(int(log_amps_per_node), t, n, out)
    return list(out)


def plan_partialTrace(num_qubits, log_amps_per_node, sorted_targets):
    ext = list(sorted_targets) + [t + num_qubits for t in sorted_targets]
    e, n = _ints(ext)
    re = (C.c_int * n)()
    lib().orc_plan_partialTrace_targets(e, n, int(log_amps_per_node), re)
    rem = (C.c_int * (2 * num_qubits))()
    lib().orc_plan_partialTrace_remaining(2 * num_qubits, e, re, n, rem)
    return list(re), list(rem)[: 2 * num_qubits - n]
