// distributed_statevector.hpp -- the public state-vector API of the drop-in host layer.
// Entry points, signatures and semantics follow the reference's src/distributed_statevector.hpp
// (:18 oneTargGate, :81 manyCtrlOneTargGate, :109 swapGate, :190 manyTargGate, :279 pauliTensor,
// :287 pauliGadget, :295 phaseGadget). What stays on the host is only the decision "local kernel or
// pairwise exchange" (qubit index vs. logNumAmpsPerNode) and the relocation planning; every loop over
// amplitudes is a CUDA kernel and every exchange an NCCL/NVLink transfer behind the C-ABI.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>

#include "communication.hpp"
#include "local_statevector.hpp"
#include "misc.hpp"
#include "states.hpp"

namespace dfsa_detail {
inline void ampToArray(const Amp& a, double out[2]) { out[0] = a.real(); out[1] = a.imag(); }
}

// A 2x2 gate on a prefix qubit mixes this shard with the partner's: amps = g[b][b]*amps + g[b][!b]*partner
static inline void dfsa_prefixOneTarg(StateVector& psi, Nat target, const AmpMatrix& gate) {
    const Nat rankTarget = target - Nat(psi.logNumAmpsPerNode);
    const Nat pairRank = Nat(flipBit(psi.rank, rankTarget));
    const Nat bit = getBit(psi.rank, rankTarget);
    double f0[2], f1[2];
    dfsa_detail::ampToArray(gate[bit][bit], f0);
    dfsa_detail::ampToArray(gate[bit][!bit], f1);
    // comm_exchangeArrays(psi.amps, psi.buffer, pairRank) + the combine loop of the reference (:28-38), fused so that
    // the shard travels in chunks while the previous chunk is being combined
    DFSA_CHECK(dfsa_xk_exchangeCombine(psi.handle, int(pairRank), f0, f1));
}

inline void distributed_statevector_oneTargGate(StateVector& psi, Nat target, AmpMatrix gate) {
    if (target < psi.logNumAmpsPerNode) local_statevector_oneTargGate(psi, target, gate);
    else dfsa_prefixOneTarg(psi, target, gate);
}

static inline void distributed_statevector_manyCtrlOneTargGate(StateVector& psi, NatArray controls, Nat target, AmpMatrix gate) {
    const Nat L = Nat(psi.logNumAmpsPerNode);
    NatArray suffixCtrls;
    Index prefixCtrlMask = 0;
    for (Nat q : controls) {
        if (q >= L) prefixCtrlMask |= Index(1) << (q - L);
        else suffixCtrls.push_back(q);
    }
    // a rank whose index fails a prefix control holds no amplitude the gate touches; its partner (same control
    // bits) fails too, so returning before any communication cannot deadlock (reference :92-93)
    if ((Index(psi.rank) & prefixCtrlMask) != prefixCtrlMask) return;

    if (target < L) { local_statevector_manyCtrlOneTargGate(psi, suffixCtrls, target, gate); return; }
    if (suffixCtrls.empty()) { dfsa_prefixOneTarg(psi, target, gate); return; }

    // prefix target, suffix controls: only the ctrl=1 sub-cube (A / 2^c amplitudes) travels
    std::sort(suffixCtrls.begin(), suffixCtrls.end());
    const Nat rankTarget = target - L;
    const Nat pairRank = Nat(flipBit(psi.rank, rankTarget));
    const Index numAmpsToMod = psi.numAmpsPerNode >> suffixCtrls.size();
    const Index allOnes = (Index(1) << suffixCtrls.size()) - 1;
    DFSA_CHECK(dfsa_k_pack(psi.handle, suffixCtrls.data(), Nat(suffixCtrls.size()), allOnes, 0));
    comm_exchangeArrays(psi.buffer, 0, psi.buffer, numAmpsToMod, numAmpsToMod, pairRank);
    const Nat bit = getBit(psi.rank, rankTarget);
    double f0[2], f1[2];
    dfsa_detail::ampToArray(gate[bit][bit], f0);
    dfsa_detail::ampToArray(gate[bit][!bit], f1);
    DFSA_CHECK(dfsa_k_combineSub(psi.handle, suffixCtrls.data(), Nat(suffixCtrls.size()), allOnes, numAmpsToMod, f0, f1));
}

static inline void distributed_statevector_swapGate(StateVector& psi, Nat qb1, Nat qb2) {
    if (qb1 > qb2) std::swap(qb1, qb2);
    const Nat L = Nat(psi.logNumAmpsPerNode);

    if (qb2 < L) { local_statevector_swapGate(psi, qb1, qb2); return; }

    if (qb1 >= L) {
        // both prefix: ranks whose two bits differ trade whole shards with the rank that has them exchanged
        const Nat alt1 = qb1 - L, alt2 = qb2 - L;
        if (getBit(psi.rank, alt1) != getBit(psi.rank, alt2)) {
            const Nat pairRank = Nat(flipBit(flipBit(psi.rank, alt1), alt2));
            comm_exchangeArrays(psi.amps, psi.buffer, pairRank);
            DFSA_CHECK(dfsa_k_copyFromBuffer(psi.handle, 0, 0, psi.numAmpsPerNode));
        }
        return;
    }

    // one suffix, one prefix qubit: the half of the shard whose qb1 bit differs from this rank's qb2 bit moves
    const Nat alt2 = qb2 - L;
    const Nat pairRank = Nat(flipBit(psi.rank, alt2));
    const Index half = psi.numAmpsPerNode / 2;
    const Nat movingBit = !getBit(psi.rank, alt2);

    if (qb1 == L - 1) {
        // that half is contiguous: no packing
        const Index offset = half * movingBit;
        comm_exchangeArrays(psi.amps, offset, psi.buffer, 0, half, pairRank);
        DFSA_CHECK(dfsa_k_copyFromBuffer(psi.handle, offset, 0, half));
        return;
    }
    // pack the moving half, trade it with the partner, unpack into the same positions (reference :160-186)
    DFSA_CHECK(dfsa_xk_swapSuffixPrefix(psi.handle, qb1, movingBit, int(pairRank)));
}

// relocation plan of manyTargGate: each prefix target (caller order) takes the lowest still-free suffix qubit
static inline NatArray dfsa_planManyTargRelocation(Nat logNumAmpsPerNode, const NatArray& targets) {
    const Index targetMask = getBitMask(targets);
    Nat nextFree = 0;
    auto advance = [&]() { while (getBit(targetMask, nextFree)) nextFree++; };
    advance();
    NatArray placed;
    placed.reserve(targets.size());
    for (Nat t : targets) {
        if (t < logNumAmpsPerNode) placed.push_back(t);
        else { placed.push_back(nextFree++); advance(); }
    }
    return placed;
}

static inline void distributed_statevector_manyTargGate(StateVector& psi, NatArray targets, AmpMatrix gate) {
    assert(targets.size() <= psi.logNumAmpsPerNode);
    const NatArray placed = dfsa_planManyTargRelocation(Nat(psi.logNumAmpsPerNode), targets);
    for (std::size_t i = 0; i < targets.size(); i++)
        if (placed[i] != targets[i]) distributed_statevector_swapGate(psi, placed[i], targets[i]);
    local_statevector_manyTargGate(psi, placed, gate);
    for (std::size_t i = 0; i < targets.size(); i++)
        if (placed[i] != targets[i]) distributed_statevector_swapGate(psi, placed[i], targets[i]);
}

static inline void distributed_statevector_pauliTensorOrGadget(StateVector& psi, const NatArray& targets, const NatArray& paulis, Amp thisAmpFac, Amp otherAmpFac) {
    assert(targets.size() == paulis.size());
    const Nat L = Nat(psi.logNumAmpsPerNode);
    Nat numY = 0, pairRank = psi.rank;
    Index maskXY = 0, maskYZ = 0;
    for (std::size_t i = 0; i < targets.size(); i++) {
        const bool isXY = (paulis[i] == X || paulis[i] == Y), isYZ = (paulis[i] == Y || paulis[i] == Z);
        if (paulis[i] == Y) numY++;
        if (isYZ) maskYZ |= Index(1) << targets[i];                 // includes prefix bits: the sign uses global indices
        if (isXY) {
            if (targets[i] >= L) pairRank = Nat(flipBit(pairRank, targets[i] - L));
            else maskXY |= Index(1) << targets[i];
        }
    }
    if (pairRank == psi.rank) {
        local_statevector_pauliTensorOrGadget_subroutine(psi, numY, maskXY, maskYZ, thisAmpFac, otherAmpFac);
        return;
    }
    double f[2], g[2];
    dfsa_detail::ampToArray(thisAmpFac, f);
    dfsa_detail::ampToArray(otherAmpFac, g);
    const int exact = (thisAmpFac == Amp(0, 0) && otherAmpFac == Amp(1, 0));
    // full-shard exchange + combine of the reference (:227-241), chunk-pipelined
    DFSA_CHECK(dfsa_xk_exchangePauliCombine(psi.handle, int(pairRank), maskXY, maskYZ, numY, f, g, exact));
}

static inline void distributed_statevector_pauliTensor(StateVector& psi, NatArray targets, NatArray paulis) {
    distributed_statevector_pauliTensorOrGadget(psi, targets, paulis, Amp(0, 0), Amp(1, 0));
}

static inline void distributed_statevector_pauliGadget(StateVector& psi, NatArray targets, NatArray paulis, Real theta) {
    distributed_statevector_pauliTensorOrGadget(psi, targets, paulis, Amp(std::cos(theta), 0), Amp(0, std::sin(theta)));
}

static inline void distributed_statevector_phaseGadget(StateVector& psi, NatArray targets, Real theta) {
    local_statevector_phaseGadget(psi, targets, theta);
}
