// local_statevector.hpp -- the rank-local state-vector operations (reference: src/local_statevector.hpp:14-153).
// Same names and argument meaning; each is one sm_100a kernel launch through the C-ABI (include/dfsa_b200.h)
// instead of an OpenMP loop. Every qubit passed here must be a suffix (rank-local) qubit.
#pragma once

#include <algorithm>
#include <cmath>

#include "states.hpp"

static inline void local_statevector_oneTargGate(StateVector& psi, Nat target, const AmpMatrix& gate) {
    std::vector<double> g = dfsaFlatten(gate);
    DFSA_CHECK(dfsa_k_ctrlOneTarg(psi.handle, nullptr, 0, target, g.data()));
}

static inline void local_statevector_manyCtrlOneTargGate(StateVector& psi, const NatArray& controls, Nat target, const AmpMatrix& gate) {
    std::vector<double> g = dfsaFlatten(gate);
    DFSA_CHECK(dfsa_k_ctrlOneTarg(psi.handle, controls.data(), Nat(controls.size()), target, g.data()));
}

static inline void local_statevector_swapGate(StateVector& psi, Nat qb1, Nat qb2) {
    DFSA_CHECK(dfsa_k_swap(psi.handle, qb1, qb2));
}

static inline void local_statevector_manyTargGate(StateVector& psi, const NatArray& targets, const AmpMatrix& gate) {
    std::vector<double> g = dfsaFlatten(gate);
    DFSA_CHECK(dfsa_k_manyTarg(psi.handle, targets.data(), Nat(targets.size()), g.data()));
}

// powI = i^(number of Y) is passed as that count; targs (the suffix X/Y qubits) is implied by maskXY
static inline void local_statevector_pauliTensorOrGadget_subroutine(StateVector& psi, Nat numY, Index maskXY, Index maskYZ, Amp thisAmpFac, Amp otherAmpFac) {
    const double f[2] = {thisAmpFac.real(), thisAmpFac.imag()}, g[2] = {otherAmpFac.real(), otherAmpFac.imag()};
    const int exact = (thisAmpFac == Amp(0, 0) && otherAmpFac == Amp(1, 0));   // pauliTensor: permutation + sign only
    DFSA_CHECK(dfsa_k_pauli(psi.handle, maskXY, maskYZ, numY, f, g, exact));
}

static inline void local_statevector_phaseGadget(StateVector& psi, const NatArray& targets, Real theta) {
    DFSA_CHECK(dfsa_k_phase(psi.handle, getBitMask(targets), theta));
}
