"""Independent dense ground truth (numpy), CHECKER ONLY.

Plays the role of the reference's own test oracle (tests/test_utilities.hpp:227-416: full 2^n x 2^n
operators applied to the gathered state, Kraus maps as sum K rho K^dagger) but is built differently:
operators act through tensor reshapes, density matrices are held as explicit 2^N x 2^N matrices.
Only usable at small sizes (<= ~12 index bits). Conventions: qubit q = bit q of
the amplitude index; gate bit i belongs to targets[i]; DM flat index = 2^N * col + row.
"""
import numpy as np

PAULI = [
    np.array([[1, 0], [0, 1]], dtype=np.complex128),
    np.array([[0, 1], [1, 0]], dtype=np.complex128),
    np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    np.array([[1, 0], [0, -1]], dtype=np.complex128),
]


def apply_to_vector(vec, targets, matrix, ctrls=()):
    """matrix (2^t x 2^t, bit i of row/col <-> targets[i]) on the sub-space where all ctrls are 1."""
    vec = np.asarray(vec, dtype=np.complex128)
    n = vec.size.bit_length() - 1
    t = len(targets)
    m = np.asarray(matrix, dtype=np.complex128).reshape([2] * (2 * t))
    psi = vec.reshape([2] * n)                       # axis a <-> qubit n-1-a
    ax = lambda q: n - 1 - q
    # matrix tensor axes: rows (bit t-1 .. bit 0), then cols (bit t-1 .. bit 0)
    targ_axes = [ax(targets[i]) for i in range(t - 1, -1, -1)]
    sel = [slice(None)] * n
    for c in ctrls:
        sel[ax(c)] = 1
    sub = psi[tuple(sel)]
    # axes of `sub` after removing the control axes
    removed = sorted(ax(c) for c in ctrls)
    def sub_axis(a):
        return a - sum(1 for r in removed if r < a)
    ta = [sub_axis(a) for a in targ_axes]
    res = np.tensordot(m, sub, axes=(list(range(t, 2 * t)), ta))     # result axes: rows..., rest...
    res = np.moveaxis(res, list(range(t)), ta)
    out = psi.copy()
    out[tuple(sel)] = res
    return out.reshape(-1)


def full_operator(num_qubits, targets, matrix, ctrls=()):
    dim = 1 << num_qubits
    cols = [apply_to_vector(np.eye(dim, dtype=np.complex128)[:, c], targets, matrix, ctrls) for c in range(dim)]
    return np.stack(cols, axis=1)


def pauli_product(paulis):
    """Kronecker product with paulis[i] acting on bit i (matches getKroneckerProductOfPaulis, test_utilities.hpp:397)."""
    prod = np.ones((1, 1), dtype=np.complex128)
    for p in paulis:
        prod = np.kron(PAULI[p], prod)
    return prod


def flat_to_matrix(flat, N):
    """Choi vector -> rho[row][col]; flat = 2^N*col + row (test_utilities.hpp:498-500)."""
    d = 1 << N
    return np.asarray(flat, dtype=np.complex128).reshape(d, d).T.copy()


def matrix_to_flat(rho):
    return np.asarray(rho, dtype=np.complex128).T.reshape(-1).copy()


def swap_matrix():
    return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def apply_op(kind, num_qubits, amps, op):
    """Ground-truth action of one op tuple on a full state (sv: vector; dm: flat Choi vector).
    Returns the new flat array (or (value) for expecPauliString, or (flat_out) for partialTrace)."""
    name = op[0]
    amps = np.asarray(amps, dtype=np.complex128)

    def unitary_on(targets, matrix, ctrls=()):
        if kind == "sv" or name.startswith("sv_"):
            return apply_to_vector(amps, targets, matrix, ctrls)
        F = full_operator(num_qubits, targets, matrix, ctrls)
        return matrix_to_flat(F @ flat_to_matrix(amps, num_qubits) @ F.conj().T)

    def kraus_on(targets, ops):
        rho = flat_to_matrix(amps, num_qubits)
        acc = np.zeros_like(rho)
        for k in ops:
            F = full_operator(num_qubits, targets, k)
            acc += F @ rho @ F.conj().T
        return matrix_to_flat(acc)

    if name == "sv_oneTargGate":
        return unitary_on([op[1]], op[2])
    if name == "sv_manyCtrlOneTargGate":
        return unitary_on([op[2]], op[3], ctrls=op[1])
    if name in ("sv_swapGate", "dm_swapGate"):
        return unitary_on([op[1], op[2]], swap_matrix())
    if name in ("sv_manyTargGate", "dm_manyTargGate"):
        return unitary_on(op[1], op[2])
    if name in ("sv_pauliTensor", "dm_pauliTensor"):
        return unitary_on(op[1], pauli_product(op[2]))
    if name in ("sv_pauliGadget", "dm_pauliGadget"):
        P = pauli_product(op[2])
        return unitary_on(op[1], np.cos(op[3]) * np.eye(P.shape[0]) + 1j * np.sin(op[3]) * P)
    if name in ("sv_phaseGadget", "dm_phaseGadget"):
        P = pauli_product([3] * len(op[1]))
        return unitary_on(op[1], np.cos(op[2]) * np.eye(P.shape[0]) + 1j * np.sin(op[2]) * P)
    if name == "dm_krausMap":
        return kraus_on(op[1], op[2])
    if name == "dm_oneQubitDephasing":
        p = op[2]
        return kraus_on([op[1]], [np.sqrt(1 - p) * PAULI[0], np.sqrt(p) * PAULI[3]])
    if name == "dm_twoQubitDephasing":
        p = op[3]
        II, IZ, ZI, ZZ = (np.kron(a, b) for a, b in ((PAULI[0], PAULI[0]), (PAULI[0], PAULI[3]), (PAULI[3], PAULI[0]), (PAULI[3], PAULI[3])))
        return kraus_on([op[1], op[2]], [np.sqrt(1 - p) * II, np.sqrt(p / 3) * IZ, np.sqrt(p / 3) * ZI, np.sqrt(p / 3) * ZZ])
    if name == "dm_oneQubitDepolarising":
        p = op[2]
        return kraus_on([op[1]], [np.sqrt(1 - p) * PAULI[0]] + [np.sqrt(p / 3) * PAULI[i] for i in (1, 2, 3)])
    if name == "dm_twoQubitDepolarising":
        # the CORRECT channel; the reference's code does not implement it
        p = op[3]
        ks = [np.sqrt(p / 15) * np.kron(PAULI[a], PAULI[b]) for a in range(4) for b in range(4)]
        ks[0] = np.sqrt(1 - p) * np.eye(4)
        return kraus_on([op[1], op[2]], ks)
    if name == "dm_damping":
        p = op[2]
        return kraus_on([op[1]], [np.array([[1, 0], [0, np.sqrt(1 - p)]]), np.array([[0, np.sqrt(p)], [0, 0]])])
    if name == "dm_expecPauliString":
        coeffs = np.asarray(op[1], dtype=np.float64).reshape(-1)
        paulis = np.asarray(op[2]).reshape(coeffs.size, num_qubits)
        H = sum(c * pauli_product(list(row)) for c, row in zip(coeffs, paulis))
        return complex(np.trace(H @ flat_to_matrix(amps, num_qubits)))
    if name == "dm_partialTrace":
        targs = sorted(op[1])
        N = num_qubits
        rho = flat_to_matrix(amps, N).reshape([2] * (2 * N))      # axes: row bits N-1..0, col bits N-1..0
        for t in reversed(targs):                                 # highest first so lower axes keep their place
            n_now = rho.ndim // 2
            rho = np.trace(rho, axis1=n_now - 1 - t, axis2=2 * n_now - 1 - t)
        d = 1 << (N - len(targs))
        return matrix_to_flat(rho.reshape(d, d))
    raise ValueError(name)
