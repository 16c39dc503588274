"""Seeded random test cases, one generator per reference API function.

Parameter distributions follow the reference's own randomized tests
(tests/tests_statevector.hpp:24-134, tests/tests_densitymatrix.hpp:26-282): random non-unitary gates
with N(0,1) entries, random distinct targets/controls, Pauli codes in {X,Y,Z} not all Z,
theta in (-pi, pi), probabilities in the channel's valid range.  Each case is an `op` tuple whose first
element is the API name; the same tuple drives the C oracle, the real reference build and the CUDA product.
"""
import numpy as np

SV_OPS = ["sv_oneTargGate", "sv_manyCtrlOneTargGate", "sv_swapGate", "sv_manyTargGate",
          "sv_pauliTensor", "sv_pauliGadget", "sv_phaseGadget"]
DM_OPS = ["dm_manyTargGate", "dm_swapGate", "dm_pauliTensor", "dm_pauliGadget", "dm_phaseGadget", "dm_krausMap",
          "dm_oneQubitDephasing", "dm_twoQubitDephasing", "dm_oneQubitDepolarising", "dm_twoQubitDepolarising",
          "dm_damping", "dm_expecPauliString", "dm_partialTrace"]
# ops whose result is a pure permutation / sign flip of the input amplitudes: compared with == (bit-exact)
EXACT_OPS = {"sv_swapGate", "dm_swapGate", "sv_pauliTensor", "dm_pauliTensor"}


def random_state(rng, nbits):
    return (rng.standard_normal(1 << nbits) + 1j * rng.standard_normal(1 << nbits)).astype(np.complex128)


def random_matrix(rng, dim):
    return (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))).astype(np.complex128)


def _unique(rng, lo, hi, n, exclude=()):
    pool = [q for q in range(lo, hi) if q not in exclude]
    return [int(x) for x in rng.permutation(pool)[:n]]


def _paulis_not_all_z(rng, n):
    p = [int(x) for x in rng.integers(1, 4, size=n)]
    if all(x == 3 for x in p):
        p[0] = int(rng.integers(1, 3))
    return p


def make_op(rng, name, nq, log_nodes, max_targets=None):
    """One random op for an nq-qubit state (sv: nq amplitude bits; dm: nq = N) spread over 2^log_nodes ranks.
    max_targets caps the size of dense gates (a 2^t x 2^t matrix is generated for t targets)."""
    if name == "sv_oneTargGate":
        return (name, int(rng.integers(0, nq)), random_matrix(rng, 2))
    if name == "sv_manyCtrlOneTargGate":
        target = int(rng.integers(0, nq))
        nc = int(rng.integers(1, nq - 1)) if nq > 2 else 1
        return (name, _unique(rng, 0, nq, nc, exclude=(target,)), target, random_matrix(rng, 2))
    if name in ("sv_swapGate", "dm_swapGate"):
        a, b = _unique(rng, 0, nq, 2)
        return (name, a, b)
    if name == "sv_manyTargGate":
        nt = int(rng.integers(1, min(nq - log_nodes, max_targets or 64) + 1))
        return (name, _unique(rng, 0, nq, nt), random_matrix(rng, 1 << nt))
    if name in ("dm_manyTargGate", "dm_krausMap"):
        max_t = min(nq - (log_nodes + 1) // 2, max_targets or 64)
        nt = int(rng.integers(1, max_t + 1))
        targets = _unique(rng, 0, nq, nt)
        if name == "dm_manyTargGate":
            return (name, targets, random_matrix(rng, 1 << nt))
        return (name, targets, [random_matrix(rng, 1 << nt) for _ in range(int(rng.integers(1, 10)))])
    if name in ("sv_pauliTensor", "dm_pauliTensor", "sv_pauliGadget", "dm_pauliGadget"):
        nt = int(rng.integers(1, nq + 1))
        targets = _unique(rng, 0, nq, nt)
        paulis = _paulis_not_all_z(rng, nt)
        if name.endswith("Tensor"):
            return (name, targets, paulis)
        return (name, targets, paulis, float(rng.uniform(-np.pi, np.pi)))
    if name in ("sv_phaseGadget", "dm_phaseGadget"):
        nt = int(rng.integers(1, nq + 1))
        return (name, _unique(rng, 0, nq, nt), float(rng.uniform(-np.pi, np.pi)))
    if name in ("dm_oneQubitDephasing", "dm_oneQubitDepolarising", "dm_damping"):
        return (name, int(rng.integers(0, nq)), float(rng.uniform(0, 0.5)))
    if name in ("dm_twoQubitDephasing", "dm_twoQubitDepolarising"):
        a, b = _unique(rng, 0, nq, 2)
        return (name, a, b, float(rng.uniform(0, 0.75)))
    if name == "dm_expecPauliString":
        nterms = int(rng.integers(1, 30))
        return (name, rng.uniform(-10, 10, size=nterms), rng.integers(0, 4, size=(nterms, nq)).astype(np.int64))
    if name == "dm_partialTrace":
        nt = int(rng.integers(1, nq - log_nodes + 1))
        return (name, _unique(rng, 0, nq, nt))
    raise ValueError(name)


def apply(state, op):
    """Run one op tuple on any backend object exposing the API-named methods."""
    return getattr(state, op[0])(*op[1:])
