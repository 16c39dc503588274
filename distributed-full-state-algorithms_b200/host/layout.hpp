// layout.hpp -- lazy logical -> physical qubit layout of a device-resident state.
//
// The reference moves a gate's prefix targets into the shard with swaps, applies the gate, and swaps them straight back
// (src/distributed_statevector.hpp:213-223); a prefix<->prefix swapGate ships whole shards between ranks (:120-137). Both
// are only RELABELLINGS of which index bit holds which qubit. Here a state remembers that relabelling instead of undoing it:
//     where[q] = the index bit ("physical" position) that currently holds logical qubit q      (identity after construction)
// Every state-vector entry point translates its qubits through `where` and then runs the reference's decision logic
// (local vs. exchange) on physical positions. manyTargGate relocates what it must and leaves it there; a swapGate of two
// qubits that both sit on rank bits just exchanges their entries -- no data moves. The layout is put back (restoreLayout)
// before anything that looks at amplitudes by index: downloads, comparisons, the density-matrix channels, expecPauliString,
// partialTrace -- so results, including the mutated input of partialTrace, are what the reference produces.
// DFSA_LAZY_LAYOUT=0 turns the laziness off (relocations are undone at once, as in the reference).
// The second half of this file is the launch plan of the deferred one-target gates (states.hpp gateQueue): with a layout that
// may change, a gate on a rank-bit qubit brings its qubit into the shard instead of exchanging whole shards (planFlush).
//
// Included at the bottom of states.hpp; uses only the C-ABI.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <utility>
#include <vector>

namespace dfsa_detail {

inline bool lazyLayoutEnabled() {
    static int on = -1;
    if (on < 0) { const char* e = std::getenv("DFSA_LAZY_LAYOUT"); on = (e && std::atoi(e) == 0) ? 0 : 1; }
    return on == 1;
}

// ---- the host-side decision of swapGate as a pure function of (rank, L = logNumAmpsPerNode, qubits): reference :109-187
struct ExchangePlan {
    enum Kind { Local = 0, Skip = 1, FullShard = 2, SubCube = 3, HalfContiguous = 4, HalfPacked = 5 };
    int kind = Local;
    Nat pairRank = 0;
    Index numAmps = 0;       // amplitudes that travel per direction
    Nat bit = 0;             // this rank's bit of the prefix qubit involved (FullShard / SubCube), or the moving bit (Half*)
};

inline ExchangePlan planSwap(Nat rank, Nat L, Nat qb1, Nat qb2) {
    ExchangePlan plan;
    if (qb1 > qb2) std::swap(qb1, qb2);
    const Index A = Index(1) << L;
    if (qb2 < L) { plan.kind = ExchangePlan::Local; return plan; }
    if (qb1 >= L) {
        // both prefix: ranks whose two bits differ trade whole shards with the rank that has them exchanged
        if (getBit(rank, qb1 - L) == getBit(rank, qb2 - L)) { plan.kind = ExchangePlan::Skip; return plan; }
        plan.kind = ExchangePlan::FullShard;
        plan.pairRank = Nat(flipBit(flipBit(rank, qb1 - L), qb2 - L));
        plan.numAmps = A;
        return plan;
    }
    // one suffix, one prefix qubit: the half of the shard whose qb1 bit differs from this rank's qb2 bit moves
    plan.pairRank = Nat(flipBit(rank, qb2 - L));
    plan.numAmps = A / 2;
    plan.bit = !getBit(rank, qb2 - L);
    plan.kind = (qb1 == L - 1) ? ExchangePlan::HalfContiguous : ExchangePlan::HalfPacked;
    return plan;
}

// swap two INDEX BITS of the distributed array (the data movement of swapGate, whatever qubits they hold)
inline void swapIndexBits(StateVector& psi, Nat p1, Nat p2) {
    if (p1 > p2) std::swap(p1, p2);
    const ExchangePlan plan = planSwap(psi.rank, Nat(psi.logNumAmpsPerNode), p1, p2);
    switch (plan.kind) {
        case ExchangePlan::Local: DFSA_CHECK(dfsa_k_swap(psi.handle, p1, p2)); return;
        case ExchangePlan::Skip: return;
        case ExchangePlan::FullShard:
            DFSA_CHECK(dfsa_x_exchange(psi.handle, DFSA_AMPS, 0, DFSA_BUFFER, 0, psi.numAmpsPerNode, int(plan.pairRank)));
            DFSA_CHECK(dfsa_k_copyFromBuffer(psi.handle, 0, 0, psi.numAmpsPerNode));      // whole shard: becomes a pointer swap
            return;
        default:
            // one suffix, one prefix bit (reference :140-186): fused over peer memory into a single pass where the ranks share
            // a node; otherwise contiguous exchange (top suffix bit) or pack / exchange / unpack, as the reference does
            DFSA_CHECK(dfsa_xk_swapSuffixPrefix(psi.handle, p1, plan.bit, int(plan.pairRank)));
            return;
    }
}
}  // namespace dfsa_detail

inline bool StateVector::layoutIsIdentity() const {
    for (Nat q = 0; q < Nat(where.size()); q++) if (where[q] != q) return false;
    return true;
}

inline NatArray StateVector::physical(const NatArray& logical) const {
    NatArray out(logical.size());
    for (std::size_t i = 0; i < logical.size(); i++) out[i] = where[logical[i]];
    return out;
}

inline Nat StateVector::numDisplacedAcrossShardBoundary() const {
    Nat n = 0;
    for (Nat q = 0; q < Nat(where.size()); q++) n += (q >= logNumAmpsPerNode && where[q] < logNumAmpsPerNode) ? 1 : 0;
    return n;
}

// The steps that put every logical qubit back on its own index bit, as a pure function of the layout (tests/test_layout_plan.py
// replays them on a numpy array). What manyTargGate leaves behind are 2-cycles (a prefix qubit sitting on a suffix bit and vice
// versa): up to four of those go in ONE relocation step; anything else (longer cycles after several gates) is sorted out
// with plain index-bit swaps.
namespace dfsa_detail {
struct RestorePlan {
    NatArray relocateSuffix, relocatePrefix;      // (suffix bit, prefix bit) pairs of the single relocation step (may be empty)
    std::vector<std::pair<Nat, Nat>> swaps;       // index-bit swaps that follow, in order
};

inline void relabel(NatArray& where, Nat posA, Nat posB) {
    for (Nat& w : where) {
        if (w == posA) w = posB;
        else if (w == posB) w = posA;
    }
}

inline RestorePlan planRestore(NatArray where, Nat L) {
    RestorePlan plan;
    const Nat n = Nat(where.size());
    for (Nat q = L; q < n && plan.relocateSuffix.size() < 4; q++) {
        const Nat p = where[q];                                    // logical prefix qubit q sits on bit p ...
        if (p < L && where[p] == q) { plan.relocateSuffix.push_back(p); plan.relocatePrefix.push_back(q); }   // ... and bit q holds logical qubit p
    }
    for (std::size_t i = 0; i < plan.relocateSuffix.size(); i++) relabel(where, plan.relocateSuffix[i], plan.relocatePrefix[i]);
    for (Nat q = 0; q < n; q++) {
        if (where[q] == q) continue;
        plan.swaps.emplace_back(where[q], q);
        relabel(where, where[q], q);
    }
    return plan;
}
}  // namespace dfsa_detail

// index bits posA and posB have just traded contents (or are declared to have, for a pure relabelling)
inline void StateVector::noteSwapped(Nat posA, Nat posB) { dfsa_detail::relabel(where, posA, posB); }

// ---- launching the deferred one-target gates (states.hpp gateQueue) --------------------------------------------------------
// A queued gate whose qubit sits on a RANK bit is not applied through a full-shard exchange (reference
// distributed_statevector.hpp:26-38, 16*A bytes per direction): the qubit is swapped into the shard, the layout remembers it,
// and the gate -- like every later gate on that qubit -- is a local one. Because the gates are deferred, the swap-in is planned
// with the whole queue in view (planFlush, a pure function; tests/test_flush_plan.py and the host-layer fuzz under tests/):
//   * gates run in call order; a run of gates whose targets all sit on suffix bits is one dfsa_k_gateSequence call;
//   * when the next gate targets a rank-bit qubit, ONE relocation step brings in that qubit and every other rank-bit qubit the
//     rest of the queue targets, as long as the suffix qubit each one evicts is not needed sooner than it is: k pairs cost
//     (1 - 2^-k) * 16A bytes per direction in one pass instead of k * 8A in k passes (dfsa_xk_relocate);
//   * the qubits evicted onto the rank bits are those whose next turn as a TARGET is furthest away (Belady's choice; controls do
//     not count, they work from rank bits too): never again in this queue first, and among those the most recently targeted one
//     -- circuits are mostly layers that visit the qubits in the same order again and again, so the most recent target is the
//     one whose next turn is furthest away. bench.py's sweep pays one log2(P)-pair relocation per layer.
namespace dfsa_detail {
// DFSA_GROUP_SWAPIN=0: one (suffix, rank) pair per relocation step, as before the queue was planned as a whole (for A/B timing)
inline Nat maxSwapInPairs() {
    static int pairs = -1;
    if (pairs < 0) { const char* e = std::getenv("DFSA_GROUP_SWAPIN"); pairs = (e && std::atoi(e) == 0) ? 1 : 4; }
    return Nat(pairs);
}

struct FlushStep {
    bool relocation = false;
    NatArray landing, prefix;              // relocation: (suffix bit, rank bit) index-bit pairs of one dfsa_xk_relocate call
    std::size_t first = 0, count = 0;      // gate run: queue entries [first, first + count)
};

inline std::vector<FlushStep> planFlush(NatArray where, Nat L, const std::vector<unsigned long long>& lastUse, const dfsa_gate1* queue, std::size_t n) {
    std::vector<FlushStep> plan;
    const Nat bits = Nat(where.size());
    const std::size_t never = ~std::size_t(0);
    std::size_t i = 0;
    while (i < n) {
        std::size_t j = i;
        while (j < n && where[queue[j].target] < L) j++;
        if (j > i) { FlushStep run; run.first = i; run.count = j - i; plan.push_back(run); }
        if (j == n) break;
        // queue[j] targets a qubit that sits on a rank bit
        std::vector<std::size_t> nextUse(bits, never);
        for (std::size_t g = n; g-- > j;) nextUse[queue[g].target] = g;
        NatArray wanted, victims;
        for (Nat q = 0; q < bits; q++) {
            if (where[q] >= L && nextUse[q] != never) wanted.push_back(q);
            if (where[q] < L) victims.push_back(q);
        }
        std::sort(wanted.begin(), wanted.end(), [&](Nat a, Nat b) { return nextUse[a] < nextUse[b]; });
        std::sort(victims.begin(), victims.end(), [&](Nat a, Nat b) {
            if (nextUse[a] != nextUse[b]) return nextUse[a] > nextUse[b];
            if (lastUse[a] != lastUse[b]) return lastUse[a] > lastUse[b];
            return where[a] > where[b];
        });
        FlushStep reloc;
        reloc.relocation = true;
        for (std::size_t w = 0; w < wanted.size() && w < victims.size() && w < maxSwapInPairs(); w++) {
            if (w > 0 && !(nextUse[victims[w]] > nextUse[wanted[w]])) break;      // it would evict a qubit that is needed sooner
            reloc.landing.push_back(where[victims[w]]);
            reloc.prefix.push_back(where[wanted[w]]);
        }
        for (std::size_t p = 0; p < reloc.landing.size(); p++) relabel(where, reloc.landing[p], reloc.prefix[p]);
        plan.push_back(reloc);
        i = j;
    }
    return plan;
}

// a queued gate (logical qubits) as the library wants it (index bits), given the layout at the time it runs
inline dfsa_gate1 physicalGate(const dfsa_gate1& logical, const NatArray& where) {
    dfsa_gate1 g = logical;
    g.target = where[logical.target];
    g.ctrlMask = 0;
    for (Nat q = 0; q < Nat(where.size()); q++) if ((logical.ctrlMask >> q) & 1ULL) g.ctrlMask |= 1ULL << where[q];
    return g;
}
}  // namespace dfsa_detail

inline void StateVector::flushGates() {
    if (gateQueue.empty()) return;
    std::vector<dfsa_gate1> queue;
    queue.swap(gateQueue);                           // (nothing below enqueues; the member is empty from here on whatever happens)
    const std::vector<dfsa_detail::FlushStep> plan = dfsa_detail::planFlush(where, Nat(logNumAmpsPerNode), lastUse, queue.data(), queue.size());
    std::vector<dfsa_gate1> run;
    for (const dfsa_detail::FlushStep& step : plan) {
        if (step.relocation) {
            DFSA_CHECK(dfsa_xk_relocate(handle, step.landing.data(), step.prefix.data(), Nat(step.landing.size())));
            for (std::size_t p = 0; p < step.landing.size(); p++) noteSwapped(step.landing[p], step.prefix[p]);
            continue;
        }
        run.clear();
        for (std::size_t g = step.first; g < step.first + step.count; g++) run.push_back(dfsa_detail::physicalGate(queue[g], where));
        DFSA_CHECK(dfsa_k_gateSequence(handle, run.data(), Nat(run.size())));
    }
}

inline void StateVector::restoreLayout() {
    flushGates();                                    // whoever asks for index order is about to look at the amplitudes
    if (layoutIsIdentity()) return;
    const dfsa_detail::RestorePlan plan = dfsa_detail::planRestore(where, Nat(logNumAmpsPerNode));
    if (!plan.relocateSuffix.empty())
        DFSA_CHECK(dfsa_xk_relocate(handle, plan.relocateSuffix.data(), plan.relocatePrefix.data(), Nat(plan.relocateSuffix.size())));
    for (const auto& sw : plan.swaps) dfsa_detail::swapIndexBits(*this, sw.first, sw.second);
    resetLayout();
}
