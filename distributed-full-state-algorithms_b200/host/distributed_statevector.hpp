// distributed_statevector.hpp -- the public state-vector API of the drop-in host layer.
// Entry points, signatures and semantics follow the reference's src/distributed_statevector.hpp
// (:18 oneTargGate, :81 manyCtrlOneTargGate, :109 swapGate, :190 manyTargGate, :279 pauliTensor,
// :287 pauliGadget, :295 phaseGadget). What stays on the host is only the decision "local kernel or
// pairwise exchange" (qubit index vs. logNumAmpsPerNode), the relocation planning, the lazy qubit layout and
// the launch plan of the deferred one-target gates (layout.hpp); every loop over amplitudes is a CUDA kernel
// and every exchange an NCCL/NVLink transfer behind the C-ABI.
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

// ---- the host-side decisions of the distributed algorithms as pure functions of (rank, L = logNumAmpsPerNode, args).
// The entry points below act on exactly these plans; tests/test_gloo_host_plans.py checks on CPU, across ranks, that
// every planned exchange is symmetric (my partner plans the same exchange with me, same size).

// (ExchangePlan and planSwap live in layout.hpp: restoring a lazy layout needs the same four swap cases)

// oneTargGate / manyCtrlOneTargGate (reference :18-39, :81-106)
inline ExchangePlan planCtrlOneTarg(Nat rank, Nat L, const NatArray& controls, Nat target, NatArray* suffixCtrlsOut = nullptr) {
    ExchangePlan plan;
    NatArray suffixCtrls;
    Index prefixCtrlMask = 0;
    for (Nat q : controls) {
        if (q >= L) prefixCtrlMask |= Index(1) << (q - L);
        else suffixCtrls.push_back(q);
    }
    std::sort(suffixCtrls.begin(), suffixCtrls.end());
    if (suffixCtrlsOut) *suffixCtrlsOut = suffixCtrls;
    // a rank whose index fails a prefix control holds no amplitude the gate touches; its partner (same control
    // bits) fails too, so skipping before any communication cannot deadlock (reference :92-93)
    if ((Index(rank) & prefixCtrlMask) != prefixCtrlMask) { plan.kind = ExchangePlan::Skip; return plan; }
    if (target < L) { plan.kind = ExchangePlan::Local; return plan; }
    plan.pairRank = Nat(flipBit(rank, target - L));
    plan.bit = getBit(rank, target - L);
    plan.kind = suffixCtrls.empty() ? ExchangePlan::FullShard : ExchangePlan::SubCube;
    plan.numAmps = (Index(1) << L) >> suffixCtrls.size();
    return plan;
}

struct PauliPlan {
    Nat pairRank = 0, numY = 0;
    Index maskXY = 0, maskYZ = 0;
};

// pauliTensor / pauliGadget (reference :244-276)
inline PauliPlan planPauli(Nat rank, Nat L, const NatArray& targets, const NatArray& paulis) {
    PauliPlan plan;
    plan.pairRank = rank;
    for (std::size_t i = 0; i < targets.size(); i++) {
        const bool isXY = (paulis[i] == X || paulis[i] == Y), isYZ = (paulis[i] == Y || paulis[i] == Z);
        if (paulis[i] == Y) plan.numY++;
        if (isYZ) plan.maskYZ |= Index(1) << targets[i];            // includes prefix bits: the sign uses global indices
        if (isXY) {
            if (targets[i] >= L) plan.pairRank = Nat(flipBit(plan.pairRank, targets[i] - L));
            else plan.maskXY |= Index(1) << targets[i];
        }
    }
    return plan;
}
}  // namespace dfsa_detail

// oneTargGate / manyCtrlOneTargGate are DEFERRED (states.hpp gateQueue) so that runs of such gates share passes over HBM; the
// queue holds logical qubits, flushGates() (layout.hpp) translates them through the layout at launch time. Controls that end up
// on rank bits stay in the mask: the library drops the gate on ranks that fail them (reference :92-93). With the lazy layout on,
// a gate whose qubit sits on a rank bit is queued like any other -- flushGates() swaps the qubit into the shard (8*A bytes per
// direction instead of the 16*A of the full-shard exchange, less when several qubits come in together) and the gate is local.
namespace dfsa_detail {
inline void enqueueOneTarg(StateVector& psi, const NatArray& controls, Nat target, const AmpMatrix& gate) {
    dfsa_gate1 g;
    const std::vector<double> flat = dfsaFlatten(gate);
    for (int i = 0; i < 8; i++) g.matrix[i] = flat[i];
    g.ctrlMask = getBitMask(controls);
    g.target = target;
    g.reserved = 0;
    psi.gateQueue.push_back(g);
    if (psi.gateQueue.size() >= 256) psi.flushGates();
}
// deferred unless fusion is off; a rank-bit qubit can only be deferred if the layout may change under it
inline bool deferOneTarg(const StateVector& psi, Nat target) {
    return StateVector::gateFusionEnabled() && (psi.where[target] < psi.logNumAmpsPerNode || lazyLayoutEnabled());
}
}  // namespace dfsa_detail

// A 2x2 gate on a prefix qubit mixes this shard with the partner's: amps = g[b][b]*amps + g[b][!b]*partner
static inline void dfsa_prefixOneTarg(StateVector& psi, Nat target, const AmpMatrix& gate) {
    psi.flushGates();
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
    psi.touch(target);
    if (dfsa_detail::deferOneTarg(psi, target)) { dfsa_detail::enqueueOneTarg(psi, {}, target, gate); return; }
    psi.flushGates();
    target = psi.where[target];                               // the index bit that holds the qubit (layout.hpp)
    if (target < psi.logNumAmpsPerNode) local_statevector_oneTargGate(psi, target, gate);
    else dfsa_prefixOneTarg(psi, target, gate);
}

static inline void distributed_statevector_manyCtrlOneTargGate(StateVector& psi, NatArray controls, Nat target, AmpMatrix gate) {
    const Nat L = Nat(psi.logNumAmpsPerNode);
    psi.touch(target);                                        // (controls do not count: a control is just as good on a rank bit)
    if (dfsa_detail::deferOneTarg(psi, target)) { dfsa_detail::enqueueOneTarg(psi, controls, target, gate); return; }
    psi.flushGates();
    controls = psi.physical(controls);
    target = psi.where[target];
    NatArray suffixCtrls;
    const dfsa_detail::ExchangePlan plan = dfsa_detail::planCtrlOneTarg(psi.rank, L, controls, target, &suffixCtrls);
    switch (plan.kind) {
        case dfsa_detail::ExchangePlan::Skip: return;
        case dfsa_detail::ExchangePlan::Local: local_statevector_manyCtrlOneTargGate(psi, suffixCtrls, target, gate); return;
        case dfsa_detail::ExchangePlan::FullShard: dfsa_prefixOneTarg(psi, target, gate); return;
        default: break;
    }
    // prefix target, suffix controls: only the ctrl=1 sub-cube (A / 2^c amplitudes) is exchanged and combined (reference
    // :43-78). Fused over peer memory where the ranks share a node, else pack / exchange / combine as the reference does.
    double f0[2], f1[2];
    dfsa_detail::ampToArray(gate[plan.bit][plan.bit], f0);
    dfsa_detail::ampToArray(gate[plan.bit][!plan.bit], f1);
    DFSA_CHECK(dfsa_xk_ctrlPrefixTarg(psi.handle, suffixCtrls.data(), Nat(suffixCtrls.size()), int(plan.pairRank), f0, f1));
}

static inline void distributed_statevector_swapGate(StateVector& psi, Nat qb1, Nat qb2) {
    const Nat L = Nat(psi.logNumAmpsPerNode);
    psi.flushGates();
    const Nat p1 = psi.where[qb1], p2 = psi.where[qb2];
    if (p1 >= L && p2 >= L && dfsa_detail::lazyLayoutEnabled()) {
        // both qubits sit on rank bits (reference :120-137 ships whole shards between the ranks whose two bits differ): the
        // swap only renames which rank holds which block, so it is recorded in the layout and no amplitude moves
        std::swap(psi.where[qb1], psi.where[qb2]);
        return;
    }
    dfsa_detail::swapIndexBits(psi, p1, p2);                  // the reference's four cases, on index bits
}

// Relocation plan of manyTargGate: every prefix target (caller order) is swapped onto a free suffix qubit for the
// duration of the gate (reference :193-210). The reference takes the LOWEST free suffix qubits; this build takes the
// HIGHEST ones: the moving half of each swap is then made of long contiguous runs (whole half-shards for the top
// qubit), so the peer reads over NVLink are fully coalesced, and the relocated targets sit above the tile's free bits,
// which is the placement the tensor-core kernel's bulk row copies want. The relocation is undone after the gate and the
// gate's arithmetic does not depend on where its bits live, so the result is identical to the reference's plan.
static inline NatArray dfsa_planManyTargRelocation(Nat logNumAmpsPerNode, const NatArray& targets) {
    const Index targetMask = getBitMask(targets);
    int nextFree = int(logNumAmpsPerNode) - 1;
    auto advance = [&]() { while (nextFree >= 0 && getBit(targetMask, Nat(nextFree))) nextFree--; };
    advance();
    NatArray placed;
    placed.reserve(targets.size());
    for (Nat t : targets) {
        if (t < logNumAmpsPerNode) placed.push_back(t);
        else { assert(nextFree >= 0); placed.push_back(Nat(nextFree--)); advance(); }
    }
    return placed;
}

// The same plan when the layout is not the identity (lazy layout, layout.hpp): `targets` are index bits, and WHICH free suffix
// bit a prefix target lands on decides how cheap the eventual restore is. In order of preference: (1) the suffix bit that
// currently holds the qubit whose home is the target's prefix bit -- the swap takes that qubit home; (2) the highest free suffix
// bit that holds its own qubit -- the swap creates a 2-cycle, which one relocation step undoes; (3) the highest free suffix bit.
// (Always taking the highest free bit lets a second prefix target evict the first one onto a rank bit: a rank relabelling that
// costs a full-shard exchange to undo -- config 3 at 8 GPUs spent 60 ms per pass on exactly that.)
static inline NatArray dfsa_planRelocationOnLayout(const NatArray& where, Nat L, const NatArray& targets) {
    const Index targetMask = getBitMask(targets);
    Index used = 0;
    auto isFree = [&](Nat s) { return s < L && !getBit(targetMask, s) && !getBit(used, s); };
    NatArray placed;
    placed.reserve(targets.size());
    for (Nat t : targets) {
        if (t < L) { placed.push_back(t); continue; }
        int choice = -1;
        if (isFree(where[t])) choice = int(where[t]);                               // (1) logical qubit t sits on a free suffix bit: bring it home
        for (int s = int(L) - 1; s >= 0 && choice < 0; s--)
            if (isFree(Nat(s)) && where[Nat(s)] == Nat(s)) choice = s;              // (2) a suffix bit holding its own qubit
        for (int s = int(L) - 1; s >= 0 && choice < 0; s--)
            if (isFree(Nat(s))) choice = s;                                         // (3) any free suffix bit
        assert(choice >= 0);
        used |= Index(1) << choice;
        placed.push_back(Nat(choice));
    }
    return placed;
}

// manyTargGate with the local step left open: `applyLocal(placed)` runs the dense-gate kernel on the (all suffix) index bits
// the targets occupy after relocation. krausMap uses it with a kernel that builds its own superoperator on the device.
template <class LocalStep>
static inline void dfsa_manyTargWithRelocation(StateVector& psi, NatArray targets, LocalStep applyLocal) {
    assert(targets.size() <= psi.logNumAmpsPerNode);
    psi.flushGates();
    targets = psi.physical(targets);
    const NatArray placed = psi.layoutIsIdentity() ? dfsa_planManyTargRelocation(Nat(psi.logNumAmpsPerNode), targets) : dfsa_planRelocationOnLayout(psi.where, Nat(psi.logNumAmpsPerNode), targets);
    // the (suffix, prefix) index-bit pairs the reference swaps one after the other before and after the local gate (:213-223);
    // the pairs are disjoint, so they commute and go in one relocation step, which is its own inverse
    NatArray landing, prefix;
    for (std::size_t i = 0; i < targets.size(); i++)
        if (placed[i] != targets[i]) { landing.push_back(placed[i]); prefix.push_back(targets[i]); }
    assert(landing.size() <= 4 && "at most 16 ranks");
    if (!landing.empty()) DFSA_CHECK(dfsa_xk_relocate(psi.handle, landing.data(), prefix.data(), Nat(landing.size())));
    applyLocal(placed);
    if (landing.empty()) return;
    if (dfsa_detail::lazyLayoutEnabled()) {
        // leave the targets where they are and remember it: the next gate on them is local, and the undo happens once,
        // when somebody needs the amplitudes in index order (StateVector::restoreLayout)
        for (std::size_t i = 0; i < landing.size(); i++) psi.noteSwapped(landing[i], prefix[i]);
        return;
    }
    DFSA_CHECK(dfsa_xk_relocate(psi.handle, landing.data(), prefix.data(), Nat(landing.size())));
}

static inline void distributed_statevector_manyTargGate(StateVector& psi, NatArray targets, AmpMatrix gate) {
    dfsa_manyTargWithRelocation(psi, targets, [&](const NatArray& placed) { local_statevector_manyTargGate(psi, placed, gate); });
}

static inline void distributed_statevector_pauliTensorOrGadget(StateVector& psi, const NatArray& targets, const NatArray& paulis, Amp thisAmpFac, Amp otherAmpFac) {
    assert(targets.size() == paulis.size());
    psi.flushGates();
    const dfsa_detail::PauliPlan plan = dfsa_detail::planPauli(psi.rank, Nat(psi.logNumAmpsPerNode), psi.physical(targets), paulis);
    if (plan.pairRank == psi.rank) {
        local_statevector_pauliTensorOrGadget_subroutine(psi, plan.numY, plan.maskXY, plan.maskYZ, thisAmpFac, otherAmpFac);
        return;
    }
    double f[2], g[2];
    dfsa_detail::ampToArray(thisAmpFac, f);
    dfsa_detail::ampToArray(otherAmpFac, g);
    const int exact = (thisAmpFac == Amp(0, 0) && otherAmpFac == Amp(1, 0));
    // full-shard exchange + combine of the reference (:227-241), fused over peer memory / chunk-pipelined
    DFSA_CHECK(dfsa_xk_exchangePauliCombine(psi.handle, int(plan.pairRank), plan.maskXY, plan.maskYZ, plan.numY, f, g, exact));
}

static inline void distributed_statevector_pauliTensor(StateVector& psi, NatArray targets, NatArray paulis) {
    distributed_statevector_pauliTensorOrGadget(psi, targets, paulis, Amp(0, 0), Amp(1, 0));
}

static inline void distributed_statevector_pauliGadget(StateVector& psi, NatArray targets, NatArray paulis, Real theta) {
    distributed_statevector_pauliTensorOrGadget(psi, targets, paulis, Amp(std::cos(theta), 0), Amp(0, std::sin(theta)));
}

static inline void distributed_statevector_phaseGadget(StateVector& psi, NatArray targets, Real theta) {
    psi.flushGates();
    local_statevector_phaseGadget(psi, psi.physical(targets), theta);
}
