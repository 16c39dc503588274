"""TEST INFRASTRUCTURE (CPU): bench.py's roofline cost model held against an independent count.

Every roofline fraction of the bench line is (algorithmic cost of an op) / (measured time); the NVLink part of that cost comes from
bench.op_cost / bench.restore_cost / bench.relocation_cost, which re-derive on the Python side what the host layer will do given
where the qubits sit. Here the same ops run through the real host layer on the CPU stand-in, which COUNTS the amplitudes each rank
pulls from other ranks' shards (hostsim_remote_amps); 16 bytes times the maximum over the ranks is what has to cross one direction of
the busiest link, whatever the kernels look like. For every op of the toy BASELINE configs 3-5 and of the sweep (default and
per-gate mode) the model's NVLink bytes must equal that count.

    DFSA_NP=8 python tests/hostsim/cost_model_check.py     -> "cost model check: ok ..." on rank 0, exit code 0
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    import standin_env
    api, standin = standin_env.install()
    import bench
    import cases
    api.comm_init()
    world, rank = api.comm_size(), api.comm_rank()
    k = world.bit_length() - 1
    lazy = bool(api.host_lib().dfsa_host_lazyLayoutEnabled())

    def max_over_ranks(v):
        box = C.c_double(float(v))
        standin.hostsim_allreduce_max(C.byref(box))
        return box.value

    def pulled(fn):
        """bytes the busiest rank pulled over the link during fn() (one direction)"""
        before = standin.hostsim_remote_amps()
        out = fn()
        return 16.0 * max_over_ranks(standin.hostsim_remote_amps() - before), out

    def nvlink_bytes(cost):
        return sum(c[1] + (c[3] if len(c) > 3 else 0.0) for c in cost)

    def restore_steps(st, bits):
        where = st.layout()
        out = (C.c_uint * (9 * len(where)))()
        hl = api.host_lib()
        hl.dfsa_host_plan_restoreLayout.restype = C.c_uint
        cnt = hl.dfsa_host_plan_restoreLayout((C.c_uint * len(where))(*where), len(where), bits - k, out)
        return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(cnt)]

    checked, nonzero = 0, 0
    rng = np.random.default_rng(11)
    # ---- configs 3-5 (the op mix of bench.config_workload at toy sizes), several passes so that the lazy layout drifts
    def workloads():
        nq = 11 + k
        ops = []
        for _ in range(10):
            ops.append(("sv_manyTargGate", [int(x) for x in rng.permutation(nq)[:5]], bench.haar(rng, 32)))
            nt = int(rng.integers(3, 7))
            ops.append(("sv_pauliGadget", [int(x) for x in rng.permutation(nq)[:nt]], [int(x) for x in rng.integers(1, 4, size=nt)], float(rng.uniform(-3, 3))))
            ops.append(("sv_phaseGadget", [int(x) for x in rng.permutation(nq)[:int(rng.integers(1, 8))]], float(rng.uniform(-3, 3))))
            ops.append(("sv_swapGate", int(rng.integers(0, nq)), int((rng.integers(1, nq) + rng.integers(0, nq)) % nq)))
        ops = [op for op in ops if op[0] != "sv_swapGate" or op[1] != op[2]]
        yield "sv", nq, ops
        N = 6 + (k + 1) // 2
        ops = []
        for q in range(N):
            ops.append(("dm_manyTargGate", [q, (q + 1) % N], bench.haar(rng, 4)))
            ops.append(("dm_oneQubitDepolarising", q, float(rng.uniform(0, 0.5))))
            ops.append(("dm_twoQubitDephasing", q, (q + 1) % N, float(rng.uniform(0, 0.5))))
            ops.append(("dm_damping", q, float(rng.uniform(0, 0.5))))
        ops.append(("dm_manyTargGate", [N - 1, 0, 2], bench.haar(rng, 8)))          # t = 3: two passes, two relocations
        ops.append(("dm_expecPauliString", rng.uniform(-1, 1, 4), rng.integers(0, 4, size=(4, N))))
        yield "dm", N, ops

    for kind, nq, ops in workloads():
        bits = nq if kind == "sv" else 2 * nq
        st = api.DeviceState(kind, nq)
        st.init_hash(3)
        for op in ops:
            cost = bench.op_cost(op, kind, nq, k, where=st.layout(), lazy=lazy)
            if op[0] in bench.INDEX_ADDRESSED_OPS:
                cost = bench.restore_cost(restore_steps(st, bits), kind, nq, k) + cost
            got, _ = pulled(lambda: cases.apply(st, op))
            want = nvlink_bytes(cost)
            assert got == want, "%s %r at %d ranks: the stand-in pulled %g bytes, the cost model says %g (layout %r)" % (kind, op[:2], world, got, want, st.layout())
            checked += 1
            nonzero += want > 0
        steps = restore_steps(st, bits)
        got, _ = pulled(st.restore_layout)
        want = nvlink_bytes(bench.restore_cost(steps, kind, nq, k))
        assert got == want, "layout restore at %d ranks: pulled %g, model %g (%r)" % (world, got, want, steps)
        checked += 1
        nonzero += want > 0
        st.close()

    # ---- the sweep, default mode: what bench's fused accounting charges (one relocation per planned step) against the count
    nq = 12 + k
    L = nq - k
    ops = bench.make_sweep(nq)
    st = api.DeviceState("sv", nq)
    st.init_hash(5)
    for layer in range(2 * nq // max(k, 1) + 3):                     # long enough for the evicted block to wrap around the register
        for op in ops:
            cases.apply(st, op)
        plan, _ = st.plan_pending_flush()
        rel, _, _ = bench.fused_plan_summary(plan, L, lambda seg: 1)
        got, _ = pulled(st.flush)
        want = sum(bench.relocation_cost(m, float(1 << L))[1] for m in rel)
        assert got == want, "sweep layer %d at %d ranks: pulled %g, fused accounting says %g (%r)" % (layer, world, got, want, rel)
        checked += 1
        nonzero += want > 0
    st.restore_layout()
    # ---- the sweep gate by gate (bench's per_gate_mode): op_cost per gate
    api.set_gate_fusion(False)
    for op in ops:
        cost = bench.op_cost(op, "sv", nq, k, where=st.layout(), lazy=lazy)
        got, _ = pulled(lambda: cases.apply(st, op))
        want = nvlink_bytes(cost)
        if op[0] == "sv_manyCtrlOneTargGate":
            # a control on a rank bit gates out half of the ranks: the model averages over ranks, the busiest rank moves all or nothing
            assert got in (0.0, want) or abs(got - want) <= want, (op[:3], got, want)
        else:
            assert got == want, "per-gate %r at %d ranks: pulled %g, model %g" % (op[:2], world, got, want)
        checked += 1
        nonzero += want > 0
    api.set_gate_fusion(True)
    st.close()
    if rank == 0:
        print("cost model check: ok at %d rank(s), %d ops compared, %d of them with NVLink traffic" % (world, checked, nonzero))
        sys.stdout.flush()
    api.comm_end()


if __name__ == "__main__":
    main()
