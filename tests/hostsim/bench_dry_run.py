"""TEST INFRASTRUCTURE (CPU): dry run of bench.py's product arm on the CPU stand-in of the C-ABI, at a toy size, at DFSA_NP ranks
(the stand-in forks them; bench.py's Job is replaced by one without torch.distributed, everything else is bench.py's own code).

Why: bench.py is what the driver runs on the GPU box, and its control flow (warm-up, timed steps taking the flush plan of every
step, fused-pass accounting, per-gate mode, end-to-end leg, the JSON line) cannot otherwise be executed where there is no GPU. The
NUMBERS of this run mean nothing (host time stamps instead of CUDA events, plain loops instead of kernels) and it is never
reported anywhere; what the test checks is that the line is produced and has the shape the contract asks for.

How: standin_env.install() (this directory, not the product) points api.py's cached library handles at
tests/hostsim/_build/libdfsa_host_on_standin.so; the host-only planners (dfsa_plan_*) still come from the real libdfsa_b200.so, which
loads without a GPU. The product itself has no switch that would let it run on anything but the CUDA library.

    [DFSA_NP=4] python tests/hostsim/bench_dry_run.py [qubits-per-rank=13]     -> the JSON line on stdout (rank 0)
"""
import ctypes as C
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)


def main():
    import standin_env
    api, standin = standin_env.install()
    PKG = standin_env.PKG
    import bench
    api.comm_init()                                        # the stand-in forks the ranks here: from now on every rank runs this script
    world, rank = api.comm_size(), api.comm_rank()

    class DryJob(bench.Job):
        """bench.Job without torch.distributed (NCCL, CUDA tensors): the reductions go through the stand-in"""

        def __init__(self, world_, rank_, local_rank_):
            self.world, self.rank, self.local_rank = world_, rank_, local_rank_
            self.k = world_.bit_length() - 1
            self.torch = self.dist = None
            self.dfsa = importlib.import_module(PKG)
            self.lib = self.dfsa.device_lib()
            self.check = self.dfsa.api.check
            self.lib.dfsa_comm_fused_active.restype = C.c_int

        def max_over_ranks(self, v):
            box = C.c_double(float(v))
            standin.hostsim_allreduce_max(C.byref(box))
            return box.value

        def broadcast_obj(self, obj):
            raise NotImplementedError("only the parity block needs it, and the dry run skips that block")

        def close(self):
            self.dfsa.comm_end()

    bench.Job = DryJob
    per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    with_configs = len(sys.argv) > 2 and sys.argv[2] == "configs"
    k = world.bit_length() - 1
    args = types.SimpleNamespace(gpus=world, steps=2, warmup=1, impl="product", qubits=per_rank + k,
                                 no_cpu_baseline=True, skip_parity=False, skip_configs=not with_configs, per_gate=False)
    if with_configs:
        # BASELINE configs 3-5 at toy sizes: the same op mix (bench.config_workload decides the real sizes), so that bench.run_config --
        # cost model, restore plans, labels, per-op tables -- runs at this rank count
        import numpy as np

        def toy_workload(name, world_, seed=7, dm_qubits=None):
            rng = np.random.default_rng(seed)
            if name == "circuit":
                nq = 11 + k
                ops = []
                for _ in range(6):
                    ops.append(("sv_manyTargGate", [int(x) for x in rng.permutation(nq)[:5]], bench.haar(rng, 32)))
                    nt = int(rng.integers(3, 7))
                    paulis = [int(x) for x in rng.integers(1, 4, size=nt)]
                    ops.append(("sv_pauliGadget", [int(x) for x in rng.permutation(nq)[:nt]], paulis, float(rng.uniform(-3, 3))))
                    ops.append(("sv_phaseGadget", [int(x) for x in rng.permutation(nq)[:int(rng.integers(1, 8))]], float(rng.uniform(-3, 3))))
                return "sv", nq, ops, "toy config 3"
            N = 6 + (k + 1) // 2
            if name == "dm":
                ops = []
                for q in range(N):
                    ops.append(("dm_manyTargGate", [q, (q + 1) % N], bench.haar(rng, 4)))
                    ops.append(("dm_oneQubitDepolarising", q, float(rng.uniform(0, 0.5))))
                    ops.append(("dm_twoQubitDephasing", q, (q + 1) % N, float(rng.uniform(0, 0.5))))
                    ops.append(("dm_damping", q, float(rng.uniform(0, 0.5))))
                return "dm", N, ops, "toy config 4"
            coeffs = rng.uniform(-10, 10, 16)
            paulis = rng.integers(0, 4, size=(16, N))
            return "dm", N, [("dm_expecPauliString", coeffs, paulis)] * 2 + [("dm_partialTrace", [0, 2]), ("dm_partialTrace", [N - 2, N - 1])], "toy config 5"

        bench.config_workload = toy_workload
    # the parity block needs the oracle and every channel of the API: only the full-size self-check part of it is kept
    bench.parity_selfcheck = lambda job: None
    bench.run_product(args, world, rank, rank)


if __name__ == "__main__":
    main()
