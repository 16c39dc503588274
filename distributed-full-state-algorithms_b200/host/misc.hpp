// misc.hpp -- host helpers of the drop-in API (reference: src/misc.hpp:16-135).
#pragma once

#include <algorithm>

#include "bit_maths.hpp"
#include "types.hpp"

// Element <row|P|col> of a Pauli string at flat density-matrix index flatInd (col bits = low N, row bits = high N).
// Closed form of the reference's per-qubit 2x2 lookup (misc.hpp:16-38): the device kernel K23 uses the same formula.
inline Amp getPauliTensorElem(const Nat* pauliCodes, Nat numQubits, Index flatInd) {
    Nat quarterTurns = 0;   // power of i
    for (Nat q = 0; q < numQubits; q++) {
        const Nat col = getBit(flatInd, q), row = getBit(flatInd, q + numQubits);
        switch (pauliCodes[q]) {
            case I: if (row != col) return Amp(0, 0); break;
            case X: if (row == col) return Amp(0, 0); break;
            case Y: if (row == col) return Amp(0, 0); quarterTurns += row ? 1 : 3; break;   // Y[1][0] = i, Y[0][1] = -i
            default: if (row != col) return Amp(0, 0); quarterTurns += row ? 2 : 0; break;  // Z[1][1] = -1
        }
    }
    switch (quarterTurns & 3u) { case 0: return Amp(1, 0); case 1: return Amp(0, 1); case 2: return Amp(-1, 0); default: return Amp(0, -1); }
}

inline bool containsOddNumY(const NatArray& paulis) {
    return (std::count(paulis.begin(), paulis.end(), Nat(Y)) & 1) != 0;
}

// sum_K conj(K) (x) K as a 4^t x 4^t matrix: row = i*d + k, col = j*d + l  (reference misc.hpp:58-81)
inline AmpMatrix getSuperoperator(const MatrixArray& krausOps) {
    const Index d = krausOps.at(0).size(), D = d * d;
    AmpMatrix super = getZeroMatrix(D);
    for (const AmpMatrix& K : krausOps)
        for (Index i = 0; i < d; i++)
            for (Index j = 0; j < d; j++) {
                const Amp cij = std::conj(K[i][j]);
                for (Index k = 0; k < d; k++)
                    for (Index l = 0; l < d; l++) super[i * d + k][j * d + l] += cij * K[k][l];
            }
    return super;
}

// partialTrace planning (reference misc.hpp:84-103): prefix targets, visited from the last to the first, take the
// highest still-free suffix qubits.
inline NatArray getReorderedAllSuffixTargets(const NatArray& targets, Nat suffixSize) {
    const Index targetMask = getBitMask(targets);
    Nat freeQubit = getNextLeftmostZeroBit(targetMask, suffixSize);
    NatArray reordered(targets.size());
    for (std::size_t q = targets.size(); q-- != 0;) {
        if (targets[q] < suffixSize) reordered[q] = targets[q];
        else { reordered[q] = freeQubit; freeQubit = getNextLeftmostZeroBit(targetMask, freeQubit); }
    }
    return reordered;
}

// where the non-traced qubits sit after those swaps, renumbered contiguously (reference misc.hpp:106-135)
inline NatArray getNonTargetedQubitOrder(Nat numAllQubits, const NatArray& originalTargets, const NatArray& reorderedTargets) {
    NatArray occupant(numAllQubits);
    for (Nat q = 0; q < numAllQubits; q++) occupant[q] = q;
    for (std::size_t q = 0; q < reorderedTargets.size(); q++)
        if (originalTargets[q] != reorderedTargets[q]) std::swap(occupant[originalTargets[q]], occupant[reorderedTargets[q]]);
    const Index tracedMask = getBitMask(reorderedTargets);
    NatArray remaining;
    for (Nat pos = 0; pos < numAllQubits; pos++)
        if (!getBit(tracedMask, pos)) remaining.push_back(occupant[pos]);
    const Index remainingMask = getBitMask(remaining);
    for (Nat& q : remaining) q = Nat(__builtin_popcountll(remainingMask & ((Index(1) << q) - 1)));   // rank among the survivors
    return remaining;
}
