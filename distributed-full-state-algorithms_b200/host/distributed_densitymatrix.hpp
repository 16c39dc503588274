// distributed_densitymatrix.hpp -- the public density-matrix API of the drop-in host layer.
// Entry points, signatures and semantics follow the reference's src/distributed_densitymatrix.hpp
// (:15 manyTargGate, :28 swapGate, :36 pauliTensor, :53 pauliGadget, :67 phaseGadget, :79 krausMap,
// :92 oneQubitDephasing, :98 twoQubitDephasing, :104 oneQubitDepolarising, :241 twoQubitDepolarising,
// :266 damping, :322 expecPauliString, :347 partialTrace). A density matrix is a Choi vector, so unitaries are
// two state-vector passes (U on the ket bits, conj(U) on the bra bits q+N); "prefix" qubits are those whose
// bra bit indexes the rank: q >= N - log2(P).
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>

#include "distributed_statevector.hpp"
#include "local_densitymatrix.hpp"

namespace dfsa_detail {
inline NatArray shifted(const NatArray& qubits, Nat by) {
    NatArray out = qubits;
    for (Nat& q : out) q += by;
    return out;
}
}

static inline void distributed_densitymatrix_manyTargGate(DensityMatrix& rho, NatArray targets, AmpMatrix gate) {
    assert(2 * targets.size() <= rho.logNumAmpsPerNode);
    if (targets.size() <= 2) {
        // U rho U^dagger in ONE pass over the shard: conj(U) (x) U applied to the 2t bits {targets, targets+N}
        // (the same 2t-target operator krausMap builds, reference :79-89) is HBM-bound for t <= 2 (<= 128 flop/amp),
        // so it costs 32*A bytes instead of the 64*A of the reference's two passes (:18-24).
        NatArray extended = targets;
        for (Nat t : targets) extended.push_back(t + rho.numQubits);
        distributed_statevector_manyTargGate(rho, extended, getSuperoperator(MatrixArray{gate}));
        return;
    }
    distributed_statevector_manyTargGate(rho, targets, gate);
    distributed_statevector_manyTargGate(rho, dfsa_detail::shifted(targets, rho.numQubits), getConjugateMatrix(gate));
}

static inline void distributed_densitymatrix_swapGate(DensityMatrix& rho, Nat qb1, Nat qb2) {
    distributed_statevector_swapGate(rho, qb1, qb2);
    distributed_statevector_swapGate(rho, qb1 + rho.numQubits, qb2 + rho.numQubits);
}

static inline void distributed_densitymatrix_pauliTensor(DensityMatrix& rho, NatArray targets, NatArray paulis) {
    distributed_statevector_pauliTensor(rho, targets, paulis);
    distributed_statevector_pauliTensor(rho, dfsa_detail::shifted(targets, rho.numQubits), paulis);
    if (containsOddNumY(paulis)) {          // conj(Y) = -Y on the bra side
        const double minusOne[2] = {-1.0, 0.0};
        DFSA_CHECK(dfsa_k_scaleAll(rho.handle, minusOne));
    }
}

static inline void distributed_densitymatrix_pauliGadget(DensityMatrix& rho, NatArray targets, NatArray paulis, Real theta) {
    distributed_statevector_pauliGadget(rho, targets, paulis, theta);
    // conj(exp(i theta P)) = exp(-i theta conj(P)), conj(P) = -P for an odd number of Y
    distributed_statevector_pauliGadget(rho, dfsa_detail::shifted(targets, rho.numQubits), paulis, containsOddNumY(paulis) ? theta : -theta);
}

static inline void distributed_densitymatrix_phaseGadget(DensityMatrix& rho, NatArray targets, Real theta) {
    distributed_statevector_phaseGadget(rho, targets, theta);
    distributed_statevector_phaseGadget(rho, dfsa_detail::shifted(targets, rho.numQubits), -theta);
}

static inline void distributed_densitymatrix_krausMap(DensityMatrix& rho, MatrixArray krausOps, NatArray targets) {
    assert(2 * targets.size() <= rho.logNumAmpsPerNode);
    NatArray extended = targets;
    for (Nat t : targets) extended.push_back(t + rho.numQubits);
    // one 2t-target gate with the superoperator sum_K conj(K) (x) K (reference :79-89); the library builds that matrix itself --
    // on the device from 4 target qubits on, where it has 65536+ entries (getSuperoperator, misc.hpp:58-81, remains available)
    std::vector<double> flat;
    for (const AmpMatrix& K : krausOps) { const std::vector<double> f = dfsaFlatten(K); flat.insert(flat.end(), f.begin(), f.end()); }
    dfsa_manyTargWithRelocation(rho, extended, [&](const NatArray& placed) {
        DFSA_CHECK(dfsa_k_krausMap(rho.handle, placed.data(), Nat(placed.size()), flat.data(), Nat(krausOps.size())));
    });
}

static inline void distributed_densitymatrix_oneQubitDephasing(DensityMatrix& rho, Nat qb, Real prob) {
    rho.restoreLayout();             // the channel kernels pair ket bit q with bra bit q + N by position
    local_densitymatrix_oneQubitDephasing(rho, qb, prob);      // the kernel handles the prefix case via the rank bit
}

static inline void distributed_densitymatrix_twoQubitDephasing(DensityMatrix& rho, Nat qb1, Nat qb2, Real prob) {
    rho.restoreLayout();             // the channel kernels pair ket bit q with bra bit q + N by position
    local_densitymatrix_twoQubitDephasing(rho, qb1, qb2, prob);
}

static inline void distributed_densitymatrix_oneQubitDepolarising(DensityMatrix& rho, Nat qb, Real prob) {
    rho.restoreLayout();             // the channel kernels pair ket bit q with bra bit q + N by position
    const Nat threshold = rho.numQubits - rho.logNumNodes;
    if (qb < threshold) { local_densitymatrix_oneQubitDepolarising(rho, qb, prob); return; }
    // the bra bit is a rank bit: the ket-bit == rank-bit halves are traded with the partner and mixed (reference :110-141);
    // one fused pass over peer memory where the ranks share a node, else pack / exchange / combine
    const Nat rankQb = qb - threshold, bit = getBit(rho.rank, rankQb);
    DFSA_CHECK(dfsa_xk_depol1Prefix(rho.handle, qb, bit, prob, int(flipBit(rho.rank, rankQb))));
}

// As in the reference, the three branches apply the reference's own formulas (SURVEY F2 explains why they are not the
// textbook channel); corrected = true applies the true channel (1-16p/15) rho + (4p/15) I (x) Tr_2 rho on all three.
// A build with -DDFSA_CORRECTED_DEPOL2_DEFAULT makes that the default (the repaired Catch2 case, tests/catch_dropin/).
#ifdef DFSA_CORRECTED_DEPOL2_DEFAULT
#define DFSA_DEPOL2_DEFAULT true
#else
#define DFSA_DEPOL2_DEFAULT false
#endif
static inline void distributed_densitymatrix_twoQubitDepolarising(DensityMatrix& rho, Nat qb1, Nat qb2, Real prob, bool corrected = DFSA_DEPOL2_DEFAULT) {
    rho.restoreLayout();
    if (qb1 > qb2) std::swap(qb1, qb2);
    const Nat N = rho.numQubits, threshold = N - rho.logNumNodes;
    if (qb2 < threshold) { local_densitymatrix_twoQubitDepolarising(rho, qb1, qb2, prob, corrected); return; }

    if (qb1 < threshold) {
        // pair (reference :146-183): one partner (the rank bit of qb2's bra); the partner's two "diagonal" amplitudes per
        // group are read over NVLink (fused) or pre-summed, packed and exchanged as an eighth of the shard (staged)
        const Nat rankQb = qb2 - threshold, bit = getBit(rho.rank, rankQb);
        DFSA_CHECK(dfsa_xk_depol2Pair(rho.handle, qb1, qb2, bit, prob, corrected ? 1 : 0, int(flipBit(rho.rank, rankQb))));
        return;
    }
    // quad (reference :187-237): the four ranks that differ in the two bra bits; one fused pass, or two sequential
    // quarter-shard exchanges
    const Nat rankQb0 = qb1 - threshold, rankQb1 = qb2 - threshold;
    const Nat bit0 = getBit(rho.rank, rankQb0), bit1 = getBit(rho.rank, rankQb1);
    DFSA_CHECK(dfsa_xk_depol2Quad(rho.handle, qb1, qb2, bit0, bit1, prob, corrected ? 1 : 0, int(flipBit(rho.rank, rankQb0)), int(flipBit(rho.rank, rankQb1))));
}

static inline void distributed_densitymatrix_damping(DensityMatrix& rho, Nat qb, Real prob) {
    rho.restoreLayout();             // the channel kernels pair ket bit q with bra bit q + N by position
    const Nat threshold = rho.numQubits - rho.logNumNodes;
    if (qb < threshold) { local_densitymatrix_damping(rho, qb, prob); return; }
    // population flows one way, from the rank holding bra bit 1 to the rank holding bra bit 0 (reference :284-317)
    const Nat rankQb = qb - threshold, bit = getBit(rho.rank, rankQb);
    DFSA_CHECK(dfsa_xk_dampingPrefix(rho.handle, qb, bit, prob, int(flipBit(rho.rank, rankQb))));
}

static inline Amp distributed_densitymatrix_expecPauliString(DensityMatrix& rho, RealArray coeffs, NatArray allPaulis) {
    assert(allPaulis.size() == coeffs.size() * rho.numQubits);
    rho.restoreLayout();
    // local sum (reference :327-340) and the all-reduce (:342): one kernel per rank publishing into the shared page, every
    // host summing the slots in rank order; comm_reduceAmp only where the ranks have no shared page
    double sum[2] = {0.0, 0.0};
    int isGlobal = 0;
    DFSA_CHECK(dfsa_kx_expecPauliString(rho.handle, coeffs.data(), Nat(coeffs.size()), allPaulis.data(), sum, &isGlobal));
    Amp value(sum[0], sum[1]);
    if (!isGlobal) comm_reduceAmp(value);
    return value;
}

static inline DensityMatrix distributed_densitymatrix_partialTrace(DensityMatrix& inRho, NatArray targets) {
    const Nat N = inRho.numQubits, L = Nat(inRho.logNumAmpsPerNode);
    assert(N - targets.size() >= inRho.logNumNodes);
    inRho.restoreLayout();           // the reference's relocation plan (and the mutated input it leaves) start from index order
    std::sort(targets.begin(), targets.end());
    const NatArray braTargets = dfsa_detail::shifted(targets, N);

    if (targets.back() + N < L) return local_densitymatrix_partialTrace(inRho, targets, braTargets);

    // some bra bits are rank bits: swap them into free suffix bits (this MUTATES inRho, as the reference does),
    // trace locally, then restore the order of the surviving qubits on the output
    NatArray extended = targets;
    extended.insert(extended.end(), braTargets.begin(), braTargets.end());
    const NatArray reordered = getReorderedAllSuffixTargets(extended, L);
    // The reference swaps the moved targets one after the other, last to first (:382-384). Each pair is (free suffix qubit,
    // prefix qubit) and no qubit occurs twice, so the swaps commute: they go as ONE relocation step (a single gather pass over
    // the 2^k shards of the rank's group) and leave inRho exactly as the reference's sequence does.
    NatArray landing, prefix;
    for (std::size_t q = reordered.size(); q-- != 0;)
        if (reordered[q] != extended[q]) { landing.push_back(reordered[q]); prefix.push_back(extended[q]); }
    assert(landing.size() <= 4 && "at most 16 ranks");
    if (!landing.empty()) DFSA_CHECK(dfsa_xk_relocate(inRho.handle, landing.data(), prefix.data(), Nat(landing.size())));

    const NatArray pairTargets(reordered.begin() + targets.size(), reordered.end());
    DensityMatrix outRho = local_densitymatrix_partialTrace(inRho, targets, pairTargets);

    NatArray remaining = getNonTargetedQubitOrder(2 * N, extended, reordered);
    for (Nat q = Nat(remaining.size()); q-- != 0;) {
        if (remaining[q] == q) continue;
        const Nat p = Nat(std::find(remaining.begin(), remaining.end(), q) - remaining.begin());
        dfsa_detail::swapIndexBits(outRho, q, p);               // data movement: outRho must END in index order
        std::swap(remaining[q], remaining[p]);
    }
    return outRho;
}
