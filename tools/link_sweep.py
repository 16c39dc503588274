#!/usr/bin/env python
"""NVLink side of the roofline, measured: achieved GB/s per direction of every fused exchange kernel
and of the plain pull microbenchmark, for one setting of DFSA_REMOTE_INFLIGHT (loads in flight per SM for kernels that read a
peer's shard). Run once per setting -- the launch geometry is fixed at first use:

  for f in 2048 4096 8192 16384; do DFSA_REMOTE_INFLIGHT=$f python -m torch.distributed.run --nproc-per-node 2 ... tools/link_sweep.py; done

One JSON line per run on rank 0 (CUDA events on the compute stream, max over ranks, best of 3 after a warm-up)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import bench  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    job = bench.Job(world, rank, int(os.environ.get("LOCAL_RANK", "0")))
    assert world > 1, "needs at least 2 GPUs"
    job.dfsa.set_gate_fusion(False)       # what is measured here are the exchange kernels of single gates: nothing may stay queued
    k = job.k
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 30 + k
    rng = np.random.default_rng(1)
    st = job.dfsa.DeviceState("sv", nq)
    st.init_hash(1)
    A = st.num_amps_per_node
    top = nq - 1

    def timed(fn, reps=3, kernel_only=False):
        """best-of-reps device time of fn (max over ranks); kernel_only: the fused exchange kernel alone, as the library timed it"""
        best = None
        for r in range(reps + 1):
            e0, e1 = job.event(), job.event()
            job.barrier()
            job.record(e0)
            fn()
            job.record(e1)
            job.barrier()
            t = job.elapsed(e0, e1)
            if kernel_only:
                ms = C.c_double()
                job.check(job.lib.dfsa_comm_last_exchange_ms(C.byref(ms)))
                t = ms.value
            t = job.max_over_ranks(t)
            if r > 0:
                best = t if best is None else min(best, t)
        return best

    out = {"n_gpus": world, "qubits": nq, "shard_GiB": 16.0 * A / 2 ** 30, "remote_inflight": int(os.environ.get("DFSA_REMOTE_INFLIGHT", "8192")),
           "fused_active": job.lib.dfsa_comm_fused_active(), "ops": {}}

    def add(label, ms, nvlink_bytes):
        out["ops"][label] = {"ms": round(ms, 3), "GBs_per_dir": round(nvlink_bytes / ms / 1e6, 1)}

    link = bench.measure_nvlink(job, st)
    out["pull"] = link
    g = bench.haar(rng, 2)
    add("oneTargGate prefix (16A)", timed(lambda: st.sv_oneTargGate(top, g)), 16.0 * A)
    add("pauliGadget X on prefix (16A)", timed(lambda: st.sv_pauliGadget([top, 3, 7], [1, 3, 2], 0.3)), 16.0 * A)
    add("manyCtrlOneTargGate prefix target, 1 suffix ctrl (8A)", timed(lambda: st.sv_manyCtrlOneTargGate([5], top, g)), 8.0 * A)
    add("swapGate suffix<->prefix (8A)", timed(lambda: st.sv_swapGate(top, 4)), 8.0 * A)
    add("swapGate top suffix<->prefix (8A)", timed(lambda: st.sv_swapGate(top, nq - k - 1)), 8.0 * A)
    g32 = bench.haar(rng, 32)

    def many():
        st.sv_manyTargGate([top, 0, 9, 13, 21], g32)
        st.restore_layout()
    t_many = timed(many)
    t_local = timed(lambda: st.sv_manyTargGate([20, 0, 9, 13, 21], g32))
    add("manyTargGate 1 prefix target: 2 relocations (2 x 8A) [local gate %.2f ms subtracted]" % t_local, t_many - t_local, 16.0 * A)
    # several rank bits at once: ONE gather pass over the 2^m shards of the group, (1 - 2^-m) * 16A bytes per direction each way
    # (also what the gate queue's launch plan uses to bring m rank-bit qubits into the shard together)
    for m in range(2, k + 1):
        def many_m():
            st.sv_manyTargGate([nq - 1 - i for i in range(m)] + [0, 9, 13, 21][: 5 - m], g32)
            st.restore_layout()
        add("manyTargGate %d prefix targets: 2 relocations of %d pairs (2 x %.0fA) [local gate subtracted]" % (m, m, (1 - 0.5 ** m) * 16), timed(many_m) - t_local, 2 * (1 - 0.5 ** m) * 16.0 * A)
    st.close()

    N = (nq + k) // 2 if (nq + k) % 2 == 0 else (nq + k - 1) // 2
    rho = job.dfsa.DeviceState("dm", N)
    rho.init_hash(2)
    Ad = rho.num_amps_per_node
    add("dm oneQubitDepolarising prefix (8A)", timed(lambda: rho.dm_oneQubitDepolarising(N - 1, 0.1)), 8.0 * Ad)
    add("dm damping prefix (8A one way)", timed(lambda: rho.dm_damping(N - 1, 0.1)), 8.0 * Ad)
    add("dm twoQubitDepolarising pair (4A)", timed(lambda: rho.dm_twoQubitDepolarising(N - 1, 2, 0.1)), 4.0 * Ad)
    if k >= 2:
        add("dm twoQubitDepolarising quad (8A)", timed(lambda: rho.dm_twoQubitDepolarising(N - 1, N - 2, 0.1)), 8.0 * Ad)
    out["dm_qubits"] = N
    rho.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    job.close()


if __name__ == "__main__":
    main()
