// local_densitymatrix.hpp -- the rank-local density-matrix operations (reference: src/local_densitymatrix.hpp:12-164),
// each one sm_100a kernel launch through the C-ABI instead of an OpenMP loop.
#pragma once

#include <algorithm>

#include "states.hpp"

static inline void local_densitymatrix_oneQubitDephasing(DensityMatrix& rho, Nat qb, Real prob) {
    DFSA_CHECK(dfsa_k_oneQubitDephasing(rho.handle, qb, prob));
}

static inline void local_densitymatrix_twoQubitDephasing(DensityMatrix& rho, Nat qb1, Nat qb2, Real prob) {
    DFSA_CHECK(dfsa_k_twoQubitDephasing(rho.handle, qb1, qb2, prob));
}

static inline void local_densitymatrix_oneQubitDepolarising(DensityMatrix& rho, Nat qb, Real prob) {
    DFSA_CHECK(dfsa_k_oneQubitDepolarising(rho.handle, qb, prob));
}

// corrected = false reproduces the reference's formulas (which are not the depolarising channel)
static inline void local_densitymatrix_twoQubitDepolarising(DensityMatrix& rho, Nat qb1, Nat qb2, Real prob, bool corrected = false) {
    DFSA_CHECK(dfsa_k_twoQubitDepolarising(rho.handle, qb1, qb2, prob, corrected ? 1 : 0));
}

static inline void local_densitymatrix_damping(DensityMatrix& rho, Nat qb, Real prob) {
    DFSA_CHECK(dfsa_k_damping(rho.handle, qb, prob));
}

// targs / pairTargs: matching ket / bra bit positions, all rank-local
static inline DensityMatrix local_densitymatrix_partialTrace(DensityMatrix& inRho, const NatArray& targs, const NatArray& pairTargs) {
    DensityMatrix outRho(inRho.numQubits - Nat(targs.size()));
    DFSA_CHECK(dfsa_k_partialTrace(inRho.handle, outRho.handle, targs.data(), pairTargs.data(), Nat(targs.size())));
    return outRho;
}
