"""ctypes binding of the drop-in host API (libdfsa_host.so -> host/*.hpp -> libdfsa_b200.so -> CUDA).

`DeviceState` mirrors the reference's StateVector / DensityMatrix plus one method per public API function,
named like the op tuples the tests use (sv_oneTargGate, dm_damping, ...). Nothing here computes: every call
goes to the native libraries, and a missing library or GPU raises DfsaError.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_dev = None
_host = None


class DfsaError(RuntimeError):
    pass


def device_lib():
    """libdfsa_b200.so (thin C-ABI over the CUDA kernels and transports)."""
    global _dev
    if _dev is None:
        path = os.path.join(_PKG, "libdfsa_b200.so")
        if not os.path.exists(path):
            raise DfsaError("libdfsa_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        _dev = C.CDLL(path, mode=C.RTLD_GLOBAL)
        _dev.dfsa_last_error.restype = C.c_char_p
        _dev.dfsa_version.restype = C.c_char_p
        _dev.dfsa_comm_transport.restype = C.c_char_p
        _dev.dfsa_stream_compute.restype = C.c_void_p
        _dev.dfsa_state_ptr.restype = C.c_void_p
        _dev.dfsa_state_ptr.argtypes = [C.c_void_p, C.c_int]
        _dev.dfsa_state_num_amps_per_node.restype = C.c_uint64
        _dev.dfsa_state_num_amps_per_node.argtypes = [C.c_void_p]
    return _dev


def host_lib():
    """libdfsa_host.so (extern "C" face of the C++ drop-in headers)."""
    global _host
    if _host is None:
        device_lib()
        path = os.path.join(_PKG, "libdfsa_host.so")
        if not os.path.exists(path):
            raise DfsaError("libdfsa_host.so is not built (run __graft_entry__.build())")
        h = C.CDLL(path)
        for name in ("dfsa_host_StateVector_new", "dfsa_host_DensityMatrix_new", "dfsa_host_dm_partialTrace", "dfsa_host_state_handle"):
            getattr(h, name).restype = C.c_void_p
        h.dfsa_host_state_numAmpsPerNode.restype = C.c_uint64
        h.dfsa_host_state_getNorm2.restype = C.c_double
        h.dfsa_host_comm_getRank.restype = C.c_uint
        h.dfsa_host_comm_getNumNodes.restype = C.c_uint
        _host = h
    return _host


def check(rc):
    if rc != 0:
        raise DfsaError("libdfsa_b200 call failed (%d): %s" % (rc, device_lib().dfsa_last_error().decode()))


def comm_init():
    """comm_init(): joins the job described by RANK/WORLD_SIZE (torchrun-style) or runs single-rank.
    Uses the non-aborting C-ABI entry so that 'no GPU' surfaces as an exception."""
    check(device_lib().dfsa_comm_init())


def comm_init_with_id(rank, size, unique_id, device=-1):
    buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
    check(device_lib().dfsa_comm_init_with_id(int(rank), int(size), buf, int(device)))


def comm_unique_id():
    buf = (C.c_char * 128)()
    check(device_lib().dfsa_comm_get_unique_id(buf))
    return bytes(buf)


def comm_end():
    check(device_lib().dfsa_comm_finalize())


def comm_rank():
    return int(device_lib().dfsa_comm_rank())


def comm_size():
    return int(device_lib().dfsa_comm_size())


def comm_synch():
    """comm_synch() of the host API: launches every state's deferred gates, then device sync + inter-rank barrier."""
    if _host is not None:
        _host.dfsa_host_comm_synch()
    else:
        check(device_lib().dfsa_comm_barrier())


def set_gate_fusion(on):
    """Deferred, fused one-target gates on (default, DFSA_FUSE_GATES) or off (one kernel per gate, launched at once)."""
    host_lib().dfsa_host_setGateFusion(int(bool(on)))


def gate_fusion_enabled():
    return bool(host_lib().dfsa_host_gateFusionEnabled())


class Gate1(C.Structure):
    """dfsa_gate1 of include/dfsa_b200.h"""
    _fields_ = [("matrix", C.c_double * 8), ("ctrlMask", C.c_uint64), ("target", C.c_uint32), ("reserved", C.c_uint32)]


def _flush_plan_buffers(num_gates, num_bits):
    return ((C.c_uint * (3 * (2 * num_gates + 1)))(), (C.c_uint * (8 * (num_gates + 1)))(), (C.c_uint * max(1, num_gates))(),
            (C.c_uint64 * max(1, num_gates))(), (C.c_uint * max(1, num_bits))())


def _decode_flush_plan(num_steps, steps, pairs, phys_target, phys_ctrl, where_out, num_bits):
    """-> (steps, layout afterwards); a step is ("relocate", [(suffix bit, rank bit), ...]) or ("gates", [(index-bit target,
    index-bit control mask), ...])"""
    out = []
    for i in range(num_steps):
        kind, count, off = steps[3 * i], steps[3 * i + 1], steps[3 * i + 2]
        if kind == 1:
            out.append(("relocate", [(int(pairs[2 * (off + p)]), int(pairs[2 * (off + p) + 1])) for p in range(count)]))
        else:
            out.append(("gates", [(int(phys_target[g]), int(phys_ctrl[g])) for g in range(off, off + count)]))
    return out, [int(where_out[q]) for q in range(num_bits)]


def plan_flush(where, log_num_amps_per_node, last_use, gates):
    """Host-only (no device): the steps StateVector::flushGates() (host/layout.hpp planFlush) takes for a queue of one-target gates
    `gates` = [(logical target, [logical controls]), ...], given the layout `where` and the per-qubit last-use stamps."""
    h = host_lib()
    h.dfsa_host_plan_flush.restype = C.c_uint
    n, bits = len(gates), len(where)
    arr = (Gate1 * max(1, n))()
    for i, (t, ctrls) in enumerate(gates):
        arr[i].target = t
        arr[i].ctrlMask = sum(1 << c for c in ctrls)
    bufs = _flush_plan_buffers(n, bits)
    cnt = h.dfsa_host_plan_flush((C.c_uint * bits)(*where), bits, int(log_num_amps_per_node), (C.c_ulonglong * bits)(*last_use), arr, n, *bufs)
    return _decode_flush_plan(cnt, *bufs, bits)


def _u32(xs):
    a = np.ascontiguousarray(np.asarray(xs).reshape(-1), dtype=np.uint32)
    n = int(a.size)
    if n == 0:
        a = np.zeros(1, dtype=np.uint32)
    return a.ctypes.data_as(C.POINTER(C.c_uint)), n        # the pointer object keeps `a` alive


def _cplx(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class DeviceState:
    """StateVector ("sv") or DensityMatrix ("dm") of the host API, amplitudes resident in HBM."""

    def __init__(self, kind, num_qubits, _ptr=None):
        h = host_lib()
        # fail with an exception (not the C++ layer's abort) when there is no device
        check(device_lib().dfsa_device_sync())
        self.kind = kind
        self.num_qubits = int(num_qubits)
        if _ptr is not None:
            self.p = C.c_void_p(_ptr)
        elif kind == "dm":
            self.p = C.c_void_p(h.dfsa_host_DensityMatrix_new(self.num_qubits))
        else:
            self.p = C.c_void_p(h.dfsa_host_StateVector_new(self.num_qubits))
        self.total_bits = (2 if kind == "dm" else 1) * self.num_qubits
        self.num_amps_per_node = int(h.dfsa_host_state_numAmpsPerNode(self.p))
        self.log_num_amps_per_node = int(h.dfsa_host_state_logNumAmpsPerNode(self.p))
        self.values = []

    def close(self):
        if getattr(self, "p", None):
            host_lib().dfsa_host_state_delete(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        """The C-ABI state. The C-ABI addresses amplitudes by index bit, so a lazily relabelled layout (host/layout.hpp) is put
        back first."""
        host_lib().dfsa_host_state_restoreLayout(self.p)
        return C.c_void_p(host_lib().dfsa_host_state_handle(self.p))

    def layout(self):
        """where[q] = index bit currently holding logical qubit q."""
        out = (C.c_uint * self.total_bits)()
        host_lib().dfsa_host_state_layout(self.p, out)
        return list(out)

    def restore_layout(self):
        host_lib().dfsa_host_state_restoreLayout(self.p)

    def flush(self):
        """Launch the one-target gates this state has deferred (host/states.hpp gateQueue)."""
        host_lib().dfsa_host_state_flushGates(self.p)

    def pending_gates(self):
        return int(host_lib().dfsa_host_state_pendingGates(self.p))

    def plan_pending_flush(self):
        """What the next flush() will do with the gates pending now: (steps, layout afterwards), steps as in plan_flush()."""
        h = host_lib()
        h.dfsa_host_state_planPendingFlush.restype = C.c_uint
        bufs = _flush_plan_buffers(self.pending_gates(), self.total_bits)
        cnt = h.dfsa_host_state_planPendingFlush(self.p, *bufs)
        return _decode_flush_plan(cnt, *bufs, self.total_bits)

    # ---- state I/O
    def set_amps(self, amps):
        a, ptr = _cplx(amps)
        assert a.size == 1 << self.total_bits
        host_lib().dfsa_host_state_setAllVecAmps(self.p, ptr)

    def get_amps(self):
        out = np.empty(1 << self.total_bits, dtype=np.complex128)
        host_lib().dfsa_host_state_getAllVecAmps(self.p, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def get_local_amps(self):
        out = np.empty(self.num_amps_per_node, dtype=np.complex128)
        check(device_lib().dfsa_state_download(self.handle, 0, C.c_uint64(0), C.c_uint64(self.num_amps_per_node), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def init_hash(self, seed):
        host_lib().dfsa_host_state_setHashAmps(self.p, C.c_ulonglong(seed))

    def init_plus(self):
        host_lib().dfsa_host_state_resetLayout(self.p)
        check(device_lib().dfsa_state_init_plus(self.handle))

    def copy_from(self, other):
        check(device_lib().dfsa_state_copy(self.handle, other.handle))

    def set_local_amps(self, amps):
        """This rank's shard only (no global array on the host)."""
        a, ptr = _cplx(amps)
        assert a.size == self.num_amps_per_node
        host_lib().dfsa_host_state_resetLayout(self.p)
        check(device_lib().dfsa_state_upload(self.handle, 0, C.c_uint64(0), C.c_uint64(a.size), ptr))

    def upload_shard_from(self, host_ptr):
        """Overwrite this rank's whole shard from (pinned) host memory. The old contents are dead, so a lazily relabelled layout
        is simply forgotten instead of being put back first."""
        host_lib().dfsa_host_state_resetLayout(self.p)
        check(device_lib().dfsa_state_upload(C.c_void_p(host_lib().dfsa_host_state_handle(self.p)), 0, C.c_uint64(0), C.c_uint64(self.num_amps_per_node), host_ptr))

    def download_shard_to(self, host_ptr):
        """This rank's shard, in index order, into (pinned) host memory."""
        check(device_lib().dfsa_state_download(self.handle, 0, C.c_uint64(0), C.c_uint64(self.num_amps_per_node), host_ptr))

    def compare(self, other):
        """On-device two-sided comparison with another state of the same shape (collective):
        (max |delta component|, number of amplitudes that differ in value, max |component| of `other`)."""
        d, ne, mr = C.c_double(), C.c_uint64(), C.c_double()
        check(device_lib().dfsa_state_compare(self.handle, other.handle, C.byref(d), C.byref(ne), C.byref(mr)))
        return d.value, int(ne.value), mr.value

    def compare_hash(self, seed):
        """The same against the state init_hash(seed) produces, regenerated on the fly."""
        d, ne, mr = C.c_double(), C.c_uint64(), C.c_double()
        check(device_lib().dfsa_state_compare_hash(self.handle, C.c_uint64(seed), C.byref(d), C.byref(ne), C.byref(mr)))
        return d.value, int(ne.value), mr.value

    def norm2(self):
        return float(host_lib().dfsa_host_state_getNorm2(self.p))

    # ---- state-vector API
    def sv_oneTargGate(self, target, gate):
        g, p = _cplx(gate)
        host_lib().dfsa_host_sv_oneTargGate(self.p, int(target), p)

    def sv_manyCtrlOneTargGate(self, ctrls, target, gate):
        g, p = _cplx(gate)
        c, n = _u32(ctrls)
        host_lib().dfsa_host_sv_manyCtrlOneTargGate(self.p, c, n, int(target), p)

    def sv_swapGate(self, q1, q2):
        host_lib().dfsa_host_sv_swapGate(self.p, int(q1), int(q2))

    def sv_manyTargGate(self, targets, gate):
        g, p = _cplx(gate)
        t, n = _u32(targets)
        host_lib().dfsa_host_sv_manyTargGate(self.p, t, n, p)

    def sv_pauliTensor(self, targets, paulis):
        t, n = _u32(targets)
        q, _ = _u32(paulis)
        host_lib().dfsa_host_sv_pauliTensor(self.p, t, q, n)

    def sv_pauliGadget(self, targets, paulis, theta):
        t, n = _u32(targets)
        q, _ = _u32(paulis)
        host_lib().dfsa_host_sv_pauliGadget(self.p, t, q, n, C.c_double(theta))

    def sv_phaseGadget(self, targets, theta):
        t, n = _u32(targets)
        host_lib().dfsa_host_sv_phaseGadget(self.p, t, n, C.c_double(theta))

    # ---- density-matrix API
    def dm_manyTargGate(self, targets, gate):
        g, p = _cplx(gate)
        t, n = _u32(targets)
        host_lib().dfsa_host_dm_manyTargGate(self.p, t, n, p)

    def dm_swapGate(self, q1, q2):
        host_lib().dfsa_host_dm_swapGate(self.p, int(q1), int(q2))

    def dm_pauliTensor(self, targets, paulis):
        t, n = _u32(targets)
        q, _ = _u32(paulis)
        host_lib().dfsa_host_dm_pauliTensor(self.p, t, q, n)

    def dm_pauliGadget(self, targets, paulis, theta):
        t, n = _u32(targets)
        q, _ = _u32(paulis)
        host_lib().dfsa_host_dm_pauliGadget(self.p, t, q, n, C.c_double(theta))

    def dm_phaseGadget(self, targets, theta):
        t, n = _u32(targets)
        host_lib().dfsa_host_dm_phaseGadget(self.p, t, n, C.c_double(theta))

    def dm_krausMap(self, targets, kraus_ops):
        k, p = _cplx(np.stack([np.asarray(m, dtype=np.complex128) for m in kraus_ops]))
        t, n = _u32(targets)
        host_lib().dfsa_host_dm_krausMap(self.p, p, len(kraus_ops), t, n)

    def dm_oneQubitDephasing(self, q, prob):
        host_lib().dfsa_host_dm_oneQubitDephasing(self.p, int(q), C.c_double(prob))

    def dm_twoQubitDephasing(self, q1, q2, prob):
        host_lib().dfsa_host_dm_twoQubitDephasing(self.p, int(q1), int(q2), C.c_double(prob))

    def dm_oneQubitDepolarising(self, q, prob):
        host_lib().dfsa_host_dm_oneQubitDepolarising(self.p, int(q), C.c_double(prob))

    def dm_twoQubitDepolarising(self, q1, q2, prob, corrected=False):
        host_lib().dfsa_host_dm_twoQubitDepolarising(self.p, int(q1), int(q2), C.c_double(prob), int(bool(corrected)))

    def dm_damping(self, q, prob):
        host_lib().dfsa_host_dm_damping(self.p, int(q), C.c_double(prob))

    def dm_expecPauliString(self, coeffs, paulis):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        p, n = _u32(paulis)
        assert n == coeffs.size * self.num_qubits
        out = (C.c_double * 2)()
        host_lib().dfsa_host_dm_expecPauliString(self.p, coeffs.ctypes.data_as(C.POINTER(C.c_double)), coeffs.size, p, out)
        v = complex(out[0], out[1])
        self.values.append(v)
        return v

    def dm_partialTrace(self, targets):
        t, n = _u32(targets)
        ptr = host_lib().dfsa_host_dm_partialTrace(self.p, t, n)
        return DeviceState("dm", self.num_qubits - n, _ptr=ptr)
