// test_utilities.hpp (repo side) -- what the reference's Catch2 cases (tests/tests_statevector.hpp:24-134,
// tests/tests_densitymatrix.hpp:26-282) need around the API under test, for the build that runs those 20 cases against
// the drop-in headers in distributed-full-state-algorithms_b200/host/ (tests/catch_dropin/build.sh).
//
// TEST INFRASTRUCTURE ONLY (category: checker; never part of the product). Two parts:
//   1. random inputs: the same draws from rand(), in the same order, as the reference's helpers
//      (tests/test_utilities.hpp:83-224), so that for a given trial the GPU build and the reference CPU build see
//      identical states, gates, targets and probabilities ("identical random states", BASELINE north_star);
//   2. an independent dense ground truth that acts on the gathered host copy by direct index arithmetic (the reference
//      builds full 2^n x 2^n operators out of Kronecker products and swap matrices, :244-416; the results are the same
//      operators applied, the construction here is this project's own and ~1000x cheaper per trial).
// The members the reference declares in src/states.hpp:27-30,57-58 and defines in its test utilities (:419-524)
// live in host/states.hpp here (device copies), so they are NOT defined in this file.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

#include "bit_maths.hpp"
#include "misc.hpp"
#include "states.hpp"
#include "types.hpp"

#include "catch_amalgamated.hpp"

const Real PI = 3.14159265358979323846;

const AmpMatrix matrI = {{1, 0}, {0, 1}};
const AmpMatrix matrX = {{0, 1}, {1, 0}};
const AmpMatrix matrY = {{0, Amp(0, -1)}, {Amp(0, 1), 0}};
const AmpMatrix matrZ = {{1, 0}, {0, -1}};

// trial count of a case: the reference's 5000 unless DFSA_CATCH_TRIALS says otherwise (many ranks on one GPU are slow)
inline int dfsaCatchTrials(int fallback) {
    const char* e = std::getenv("DFSA_CATCH_TRIALS");
    return (e && std::atoi(e) > 0) ? std::atoi(e) : fallback;
}

inline void rootNodePrint(const std::string& msg) {
    if (comm_getRank() == 0) std::cout << msg << std::endl;
}

// ---------------------------------------------------------------------------------------------- random inputs
// one rand() per integer / real, two per amplitude (Box-Muller: radius from the first draw, angle from the second)

namespace dfsa_test {
inline Real unitDraw() { return std::rand() / Real(RAND_MAX); }
}

inline Nat getRandomNat(Nat minInclusive, Nat maxExclusive) {
    assert(maxExclusive > minInclusive);
    return minInclusive + Nat(std::rand() % (maxExclusive - minInclusive));
}

inline Real getRandomReal(Real minInclusive, Real maxExclusive) {
    return minInclusive + dfsa_test::unitDraw() * (maxExclusive - minInclusive);
}

inline RealArray getRandomRealArray(Real minInclusive, Real maxExclusive, Nat numElems) {
    RealArray out;
    out.reserve(numElems);
    while (out.size() < numElems) out.push_back(getRandomReal(minInclusive, maxExclusive));
    return out;
}

inline Amp getRandomAmp() {
    const Real u = dfsa_test::unitDraw(), v = dfsa_test::unitDraw();
    const Real radius = std::sqrt(-2 * std::log(u)), angle = 2 * 3.14159265 * v;
    return Amp(radius * std::cos(angle), radius * std::sin(angle));
}

inline AmpArray getRandomArray(Index dim) {
    AmpArray out(dim);
    for (Amp& a : out) a = getRandomAmp();
    return out;
}

inline AmpMatrix getRandomMatrix(Index dim) {
    AmpMatrix out(dim);
    for (AmpArray& row : out) row = getRandomArray(dim);      // row-major draw order
    return out;
}

inline MatrixArray getRandomMatrices(Index dim, Nat numMatrices) {
    MatrixArray out;
    for (Nat m = 0; m < numMatrices; m++) out.push_back(getRandomMatrix(dim));
    return out;
}

inline NatArray getRandomNatArray(Nat minIncl, Nat maxExcl, Nat numElem) {
    NatArray out(numElem);
    for (Nat& x : out) x = getRandomNat(minIncl, maxExcl);
    return out;
}

// the first numElem entries of [minIncl, maxExcl) after 10 * size random transpositions
inline NatArray getRandomUniqueNatArray(Nat minIncl, Nat maxExcl, Nat numElem) {
    const Nat size = maxExcl - minIncl;
    assert(numElem >= 1 && numElem <= size);
    NatArray pool(size);
    for (Nat k = 0; k < size; k++) pool[k] = minIncl + k;
    for (Nat rep = 0; rep < 10 * size; rep++) {
        const Nat a = getRandomNat(0, size);
        const Nat b = getRandomNat(0, size);
        std::swap(pool[a], pool[b]);
    }
    pool.resize(numElem);
    return pool;
}

// ... and `exclude`, if drawn, replaced by the smallest value not drawn
inline NatArray getRandomUniqueNatArray(Nat minIncl, Nat maxExcl, Nat numElem, Nat exclude) {
    assert(maxExcl - minIncl > numElem);
    NatArray picked = getRandomUniqueNatArray(minIncl, maxExcl, numElem);
    auto hit = std::find(picked.begin(), picked.end(), exclude);
    if (hit != picked.end()) {
        Nat spare = minIncl;
        while (std::find(picked.begin(), picked.end(), spare) != picked.end()) spare++;
        *hit = spare;
    }
    return picked;
}

inline void ensureNotAllPauliZ(NatArray& paulis) {
    if (std::all_of(paulis.begin(), paulis.end(), [](Nat p) { return p == 3; })) paulis[0] = getRandomNat(1, 3);
}

// ---------------------------------------------------------------------------------------------- dense ground truth

// a (x) b with b on the low bits
inline AmpMatrix getKroneckerProduct(const AmpMatrix& a, const AmpMatrix& b) { return a % b; }

// element [r][c] = product over i of P_i[bit i of r][bit i of c]: paulis[0] acts on the lowest bit
inline AmpMatrix getKroneckerProductOfPaulis(const NatArray& paulis) {
    const AmpMatrix* table[4] = {&matrI, &matrX, &matrY, &matrZ};
    const Index dim = Index(1) << paulis.size();
    AmpMatrix out = getZeroMatrix(dim);
    for (Index r = 0; r < dim; r++)
        for (Index c = 0; c < dim; c++) {
            Amp e(1, 0);
            for (std::size_t i = 0; i < paulis.size() && e != Amp(0, 0); i++) {
                assert(paulis[i] <= 3);
                e *= (*table[paulis[i]])[(r >> i) & 1][(c >> i) & 1];
            }
            out[r][c] = e;
        }
    return out;
}

// exp(i angle P) for P^2 = 1
inline AmpMatrix getExponentialOfPauliTensor(Real angle, const AmpMatrix& pauliTensor) {
    AmpMatrix out = Amp(0, std::sin(angle)) * pauliTensor;
    for (std::size_t d = 0; d < out.size(); d++) out[d][d] += std::cos(angle);
    return out;
}

namespace dfsa_test {
// v <- (controlled gate) v for a vector addressed through `at(i)`: for every assignment of the other bits with all
// controls set, the 2^t amplitudes spanned by the targets (gate bit k <-> targs[k]) are multiplied by the gate
template <class At>
void applyGate(Index dim, const NatArray& ctrls, const NatArray& targs, const AmpMatrix& gate, At at) {
    const Index gdim = Index(1) << targs.size();
    assert(gate.size() == gdim);
    Index ctrlMask = 0, targMask = 0;
    for (Nat q : ctrls) ctrlMask |= Index(1) << q;
    for (Nat q : targs) targMask |= Index(1) << q;
    assert((ctrlMask & targMask) == 0);
    AmpArray in(gdim);
    std::vector<Index> where(gdim);
    for (Index base = 0; base < dim; base++) {
        if ((base & targMask) != 0 || (base & ctrlMask) != ctrlMask) continue;
        for (Index k = 0; k < gdim; k++) {
            Index idx = base;
            for (std::size_t b = 0; b < targs.size(); b++) idx |= ((k >> b) & 1ULL) << targs[b];
            where[k] = idx;
            in[k] = at(idx);
        }
        for (Index r = 0; r < gdim; r++) {
            Amp acc(0, 0);
            for (Index c = 0; c < gdim; c++) acc += gate[r][c] * in[c];
            at(where[r]) = acc;
        }
    }
}
}  // namespace dfsa_test

inline void applyGateToLocalState(AmpArray& state, NatArray ctrls, NatArray targs, AmpMatrix gateMatr) {
    dfsa_test::applyGate(state.size(), ctrls, targs, gateMatr, [&](Index i) -> Amp& { return state[i]; });
}

// rho <- G rho G^dagger: G on the row index of every column, conj(G) on the column index of every row
inline void applyGateToLocalState(AmpMatrix& state, NatArray ctrls, NatArray targs, AmpMatrix gateMatr) {
    const Index dim = state.size();
    for (Index c = 0; c < dim; c++)
        dfsa_test::applyGate(dim, ctrls, targs, gateMatr, [&](Index r) -> Amp& { return state[r][c]; });
    const AmpMatrix conj = getConjugateMatrix(gateMatr);
    for (Index r = 0; r < dim; r++)
        dfsa_test::applyGate(dim, ctrls, targs, conj, [&](Index c) -> Amp& { return state[r][c]; });
}

// rho <- sum_K K rho K^dagger
inline void applyKrausMapToLocalState(AmpMatrix& state, NatArray targets, MatrixArray krausOps) {
    const AmpMatrix before = state;
    state = getZeroMatrix(before.size());
    for (const AmpMatrix& K : krausOps) {
        AmpMatrix term = before;
        applyGateToLocalState(term, {}, targets, K);
        state = state + term;
    }
}
