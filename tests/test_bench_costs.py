"""CPU test of bench.py's roofline accounting (SURVEY 8d figures): the algorithmic cost of every op family, with and without the
lazy layout, and of a layout restore -- the numbers every `roofline_frac` of the bench line is a ratio of."""
import bench

PEAKS = {"hbm_GBs": 6400.0, "nvlink_GBs_per_dir": 640.0, "nvlink_one_way_GBs": 800.0, "fp64_TFLOPs": 32.0}


def test_state_vector_costs():
    A = float(1 << 32)
    assert bench.op_cost(("sv_oneTargGate", 3, None), "sv", 32, 0) == [(32 * A, 0, 0)]
    assert bench.op_cost(("sv_manyCtrlOneTargGate", [1, 40], 5, None), "sv", 34, 2) == [(16 * A, 0, 0)]            # one suffix control halves the work
    assert bench.op_cost(("sv_oneTargGate", 33, None), "sv", 34, 2) == [(48 * A, 16 * A, 0)]                        # full-shard exchange
    assert bench.op_cost(("sv_swapGate", 33, 32), "sv", 34, 2) == [(32 * A, 16 * A, 0)]
    assert bench.op_cost(("sv_swapGate", 33, 32), "sv", 34, 2, lazy=True) == []                                       # rank relabelling
    assert bench.op_cost(("sv_swapGate", 33, 3), "sv", 34, 2) == [(24 * A, 8 * A, 0)]
    # a dense gate with one prefix target: two relocations as the reference does, one with the lazy layout
    many = ("sv_manyTargGate", [33, 0, 1, 2, 3], None)
    assert len(bench.op_cost(many, "sv", 34, 2)) == 3 and len(bench.op_cost(many, "sv", 34, 2, lazy=True)) == 2
    # ... and none if the layout already holds that qubit on a suffix bit
    where = list(range(34))
    where[33], where[20] = 20, 33
    assert len(bench.op_cost(many, "sv", 34, 2, where=where, lazy=True)) == 1
    assert bench.op_cost(("sv_oneTargGate", 20, None), "sv", 34, 2, where=where) == [(48 * A, 16 * A, 0)]           # logical 20 now sits on a rank bit


def test_density_matrix_costs_and_bounds():
    A = float(1 << 29)                                             # 16 qubits on 8 ranks
    assert bench.op_cost(("dm_twoQubitDephasing", 1, 2, 0.1), "dm", 16, 3) == [(28 * A, 0, 0)]
    assert bench.op_cost(("dm_damping", 15, 0.1), "dm", 16, 3) == [(40 * A, 0, 0, 8 * A)]                             # one-way transfer
    one_pass = bench.op_cost(("dm_manyTargGate", [0, 1], None), "dm", 16, 3)
    assert one_pass == [(32 * A, 0.0, 6.0 * 16 * A)]                                                                  # U (x) conj(U) in ONE pass
    # bounds: the slowest of the three resources per phase, summed over phases
    ms = bench.bound_ms([(32 * A, 8 * A, 0.0), (0.0, 0.0, 3.2e13)], PEAKS)
    assert abs(ms - (max(32 * A / 6.4e12, 8 * A / 6.4e11) + 1.0) * 1e3) < 1e-9
    assert abs(bench.bound_ms([(40 * A, 0, 0, 8 * A)], PEAKS) - max(40 * A / 6.4e12, 8 * A / 8e11) * 1e3) < 1e-9
    # restoring a layout: all relocation pairs in one step, then index-bit swaps by kind
    cost = bench.restore_cost([(0, 28, 30), (0, 27, 31), (1, 3, 5), (1, 4, 30), (1, 30, 31)], "dm", 16, 3)
    assert cost == [(32 * A, 0.75 * 16 * A, 0.0), (16 * A, 0.0, 0.0), (24 * A, 8 * A, 0.0), (32 * A, 16 * A, 0.0)]
    assert bench.restore_cost([], "dm", 16, 3) == []


def test_inverse_ops_and_workloads_are_deterministic():
    ops = bench.make_sweep(12)
    assert [o[0] for o in bench.inverse_ops(ops)] == [o[0] for o in reversed(ops)]
    for name in ("circuit", "dm", "expec"):
        a = bench.config_workload(name, 8)
        b = bench.config_workload(name, 8)
        assert a[1] == b[1] and len(a[2]) == len(b[2]) and [o[0] for o in a[2]] == [o[0] for o in b[2]]


def test_fused_step_accounting_is_a_pure_function_of_the_flush_plan():
    A = float(1 << 32)
    assert bench.relocation_cost(1, A) == (32 * A, 8 * A, 0.0)
    assert bench.relocation_cost(3, A) == (32 * A, 14 * A, 0.0)
    plan = [("gates", [(0, 0), (1, 0b100), (2, 1 << 40)]), ("relocate", [(31, 32), (30, 33)]), ("gates", [(31, 0)])]
    seen = []

    def count_passes(segment):
        seen.append(segment)
        return 2 if len(segment) > 1 else 1

    rel, passes, summary = bench.fused_plan_summary(plan, 32, count_passes)
    assert rel == [2] and passes == 3 and len(summary) == 3
    assert seen == [[(0, []), (1, [2]), (2, [])], [(31, [])]]              # controls on rank bits do not reach the pass planner
    # the accounting never raises: a broken plan entry becomes an error string, the decodable steps still count
    class Lib:
        @staticmethod
        def dfsa_plan_gateSequence(*a):
            raise RuntimeError("no library here")
    out = bench.fused_accounting([plan, "plan_pending_flush failed: boom"], [("sv_oneTargGate", 0, None)], Lib, 32, lambda rc: None)
    assert out["error"] and out["passes"] == [] and out["relocation_pairs"] == []
    out = bench.fused_accounting(None, [("sv_oneTargGate", 0, None), ("sv_manyCtrlOneTargGate", [1, 2], 0, None)], Lib, 32, lambda rc: None)
    assert out["pairs"] == A / 2 + A / 8 and out["error"] is None


def test_numa_binding_helper_is_best_effort():
    assert bench.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and bench.parse_cpulist("") == set()
    import os
    before = os.sched_getaffinity(0)
    res = bench.bind_to_gpu_numa_node(0)                 # no nvidia-smi / no GPU here: reports why, changes nothing, raises nothing
    assert isinstance(res, dict) and "bound" in res
    if not res["bound"]:
        assert os.sched_getaffinity(0) == before and res["why"]
    else:
        os.sched_setaffinity(0, before)
