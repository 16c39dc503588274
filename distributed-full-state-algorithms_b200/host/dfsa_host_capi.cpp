// dfsa_host_capi.cpp -> libdfsa_host.so : extern "C" face of the drop-in host C++ API (host/*.hpp), so that
// non-C++ callers (the Python parity tests and bench.py via ctypes, or any FFI) drive exactly the code path a
// C++ user of distributed_statevector.hpp / distributed_densitymatrix.hpp gets:
//     ctypes -> these wrappers -> host dispatch (local vs. exchange, relocation) -> C-ABI of libdfsa_b200.so -> CUDA.
// One function per public entry point of the reference API, same argument order; vectors become (pointer, count),
// matrices row-major interleaved (re,im) doubles. Handles are StateVector* / DensityMatrix*.
#include <cstring>

#include "distributed_densitymatrix.hpp"
#include "distributed_statevector.hpp"

namespace {
AmpMatrix toMatrix(const double* flat, Index dim) {
    AmpMatrix m = getZeroMatrix(dim);
    for (Index r = 0; r < dim; r++)
        for (Index c = 0; c < dim; c++) m[r][c] = Amp(flat[2 * (r * dim + c)], flat[2 * (r * dim + c) + 1]);
    return m;
}
NatArray toNats(const unsigned* p, unsigned n) { return NatArray(p, p + n); }
StateVector& sv(void* h) { return *static_cast<StateVector*>(h); }
DensityMatrix& dm(void* h) { return *static_cast<DensityMatrix*>(h); }
}  // namespace

extern "C" {

// ---- environment (communication.hpp)
void dfsa_host_comm_init() { comm_init(); }
void dfsa_host_comm_end() { comm_end(); }
unsigned dfsa_host_comm_getRank() { return comm_getRank(); }
unsigned dfsa_host_comm_getNumNodes() { return comm_getNumNodes(); }
void dfsa_host_comm_synch() { comm_synch(); }

// ---- states (states.hpp)
void* dfsa_host_StateVector_new(unsigned numQubits) { return new StateVector(numQubits); }
void* dfsa_host_DensityMatrix_new(unsigned numQubits) { return new DensityMatrix(numQubits); }
void dfsa_host_state_delete(void* h) { delete static_cast<StateVector*>(h); }
void* dfsa_host_state_handle(void* h) { return sv(h).handle; }
unsigned dfsa_host_state_numQubits(void* h) { return sv(h).numQubits; }
unsigned long long dfsa_host_state_numAmpsPerNode(void* h) { return sv(h).numAmpsPerNode; }
unsigned dfsa_host_state_logNumAmpsPerNode(void* h) { return unsigned(sv(h).logNumAmpsPerNode); }
void dfsa_host_state_getAllVecAmps(void* h, double* out) {
    AmpArray all = sv(h).getAllVecAmps();
    std::memcpy(out, all.data(), all.size() * sizeof(Amp));
}
void dfsa_host_state_setAllVecAmps(void* h, const double* in) {
    StateVector& s = sv(h);
    const Amp* p = reinterpret_cast<const Amp*>(in);
    s.setAllVecAmps(AmpArray(p, p + Index(s.numNodes) * s.numAmpsPerNode));
}
void dfsa_host_state_setHashAmps(void* h, unsigned long long seed) { sv(h).setHashAmps(seed); }
// lazy layout (layout.hpp): out[q] = index bit holding logical qubit q; restore = put every qubit back on its own bit
void dfsa_host_state_layout(void* h, unsigned* out) { for (std::size_t q = 0; q < sv(h).where.size(); q++) out[q] = sv(h).where[q]; }
void dfsa_host_state_restoreLayout(void* h) { sv(h).restoreLayout(); }
void dfsa_host_state_resetLayout(void* h) { sv(h).resetLayout(); }
int dfsa_host_lazyLayoutEnabled() { return dfsa_detail::lazyLayoutEnabled() ? 1 : 0; }
// deferred one-target gates (states.hpp gateQueue): launch what is pending / how many are pending / switch deferral on and off
void dfsa_host_state_flushGates(void* h) { sv(h).flushGates(); }
unsigned dfsa_host_state_pendingGates(void* h) { return unsigned(sv(h).gateQueue.size()); }
int dfsa_host_gateFusionEnabled() { return StateVector::gateFusionEnabled() ? 1 : 0; }
void dfsa_host_setGateFusion(int on) { StateVector::flushAllStates(); StateVector::gateFusionEnabled() = (on != 0); }
double dfsa_host_state_getNorm2(void* h) { return sv(h).getNorm2(); }

// ---- state-vector API (distributed_statevector.hpp)
void dfsa_host_sv_oneTargGate(void* h, unsigned target, const double* gate) {
    distributed_statevector_oneTargGate(sv(h), target, toMatrix(gate, 2));
}
void dfsa_host_sv_manyCtrlOneTargGate(void* h, const unsigned* ctrls, unsigned numCtrls, unsigned target, const double* gate) {
    distributed_statevector_manyCtrlOneTargGate(sv(h), toNats(ctrls, numCtrls), target, toMatrix(gate, 2));
}
void dfsa_host_sv_swapGate(void* h, unsigned qb1, unsigned qb2) { distributed_statevector_swapGate(sv(h), qb1, qb2); }
void dfsa_host_sv_manyTargGate(void* h, const unsigned* targets, unsigned numTargets, const double* gate) {
    distributed_statevector_manyTargGate(sv(h), toNats(targets, numTargets), toMatrix(gate, powerOf2(numTargets)));
}
void dfsa_host_sv_pauliTensor(void* h, const unsigned* targets, const unsigned* paulis, unsigned n) {
    distributed_statevector_pauliTensor(sv(h), toNats(targets, n), toNats(paulis, n));
}
void dfsa_host_sv_pauliGadget(void* h, const unsigned* targets, const unsigned* paulis, unsigned n, double theta) {
    distributed_statevector_pauliGadget(sv(h), toNats(targets, n), toNats(paulis, n), theta);
}
void dfsa_host_sv_phaseGadget(void* h, const unsigned* targets, unsigned n, double theta) {
    distributed_statevector_phaseGadget(sv(h), toNats(targets, n), theta);
}

// ---- density-matrix API (distributed_densitymatrix.hpp)
void dfsa_host_dm_manyTargGate(void* h, const unsigned* targets, unsigned numTargets, const double* gate) {
    distributed_densitymatrix_manyTargGate(dm(h), toNats(targets, numTargets), toMatrix(gate, powerOf2(numTargets)));
}
void dfsa_host_dm_swapGate(void* h, unsigned qb1, unsigned qb2) { distributed_densitymatrix_swapGate(dm(h), qb1, qb2); }
void dfsa_host_dm_pauliTensor(void* h, const unsigned* targets, const unsigned* paulis, unsigned n) {
    distributed_densitymatrix_pauliTensor(dm(h), toNats(targets, n), toNats(paulis, n));
}
void dfsa_host_dm_pauliGadget(void* h, const unsigned* targets, const unsigned* paulis, unsigned n, double theta) {
    distributed_densitymatrix_pauliGadget(dm(h), toNats(targets, n), toNats(paulis, n), theta);
}
void dfsa_host_dm_phaseGadget(void* h, const unsigned* targets, unsigned n, double theta) {
    distributed_densitymatrix_phaseGadget(dm(h), toNats(targets, n), theta);
}
void dfsa_host_dm_krausMap(void* h, const double* krausOps, unsigned numOps, const unsigned* targets, unsigned numTargets) {
    const Index d = powerOf2(numTargets);
    MatrixArray ops;
    for (unsigned o = 0; o < numOps; o++) ops.push_back(toMatrix(krausOps + 2 * d * d * o, d));
    distributed_densitymatrix_krausMap(dm(h), ops, toNats(targets, numTargets));
}
void dfsa_host_dm_oneQubitDephasing(void* h, unsigned qb, double prob) { distributed_densitymatrix_oneQubitDephasing(dm(h), qb, prob); }
void dfsa_host_dm_twoQubitDephasing(void* h, unsigned qb1, unsigned qb2, double prob) { distributed_densitymatrix_twoQubitDephasing(dm(h), qb1, qb2, prob); }
void dfsa_host_dm_oneQubitDepolarising(void* h, unsigned qb, double prob) { distributed_densitymatrix_oneQubitDepolarising(dm(h), qb, prob); }
void dfsa_host_dm_twoQubitDepolarising(void* h, unsigned qb1, unsigned qb2, double prob, int corrected) {
    distributed_densitymatrix_twoQubitDepolarising(dm(h), qb1, qb2, prob, corrected != 0);
}
void dfsa_host_dm_damping(void* h, unsigned qb, double prob) { distributed_densitymatrix_damping(dm(h), qb, prob); }
void dfsa_host_dm_expecPauliString(void* h, const double* coeffs, unsigned numTerms, const unsigned* paulis, double* outReIm) {
    DensityMatrix& rho = dm(h);
    Amp v = distributed_densitymatrix_expecPauliString(rho, RealArray(coeffs, coeffs + numTerms), toNats(paulis, numTerms * rho.numQubits));
    outReIm[0] = v.real();
    outReIm[1] = v.imag();
}
void* dfsa_host_dm_partialTrace(void* h, const unsigned* targets, unsigned numTargets) {
    return new DensityMatrix(distributed_densitymatrix_partialTrace(dm(h), toNats(targets, numTargets)));
}

// ---- host-side plans (pure functions, no device needed): what each rank would exchange for an op
// out = {kind, pairRank, numAmps, bit}
void dfsa_host_plan_ctrlOneTarg(unsigned rank, unsigned L, const unsigned* ctrls, unsigned numCtrls, unsigned target, unsigned long long* out) {
    dfsa_detail::ExchangePlan p = dfsa_detail::planCtrlOneTarg(rank, L, toNats(ctrls, numCtrls), target);
    out[0] = p.kind; out[1] = p.pairRank; out[2] = p.numAmps; out[3] = p.bit;
}
void dfsa_host_plan_swap(unsigned rank, unsigned L, unsigned qb1, unsigned qb2, unsigned long long* out) {
    dfsa_detail::ExchangePlan p = dfsa_detail::planSwap(rank, L, qb1, qb2);
    out[0] = p.kind; out[1] = p.pairRank; out[2] = p.numAmps; out[3] = p.bit;
}
// out = {pairRank, numY, maskXY, maskYZ}
void dfsa_host_plan_pauli(unsigned rank, unsigned L, const unsigned* targets, const unsigned* paulis, unsigned n, unsigned long long* out) {
    dfsa_detail::PauliPlan p = dfsa_detail::planPauli(rank, L, toNats(targets, n), toNats(paulis, n));
    out[0] = p.pairRank; out[1] = p.numY; out[2] = p.maskXY; out[3] = p.maskYZ;
}
void dfsa_host_plan_manyTarg(unsigned L, const unsigned* targets, unsigned n, unsigned* placedOut) {
    NatArray placed = dfsa_planManyTargRelocation(L, toNats(targets, n));
    for (unsigned i = 0; i < n; i++) placedOut[i] = placed[i];
}
// lazy layout: where the (index-bit) targets of a dense gate land, given the current layout
void dfsa_host_plan_relocationOnLayout(const unsigned* where, unsigned n, unsigned L, const unsigned* targets, unsigned nt, unsigned* placedOut) {
    NatArray placed = dfsa_planRelocationOnLayout(NatArray(where, where + n), L, toNats(targets, nt));
    for (unsigned i = 0; i < nt; i++) placedOut[i] = placed[i];
}
// lazy layout: the steps that restore index order. out = numSteps x {kind (0 = relocation pair, 1 = index-bit swap), a, b}
unsigned dfsa_host_plan_restoreLayout(const unsigned* where, unsigned n, unsigned L, unsigned* out) {
    dfsa_detail::RestorePlan plan = dfsa_detail::planRestore(NatArray(where, where + n), L);
    unsigned steps = 0;
    for (std::size_t i = 0; i < plan.relocateSuffix.size(); i++, steps++) { out[3 * steps] = 0; out[3 * steps + 1] = plan.relocateSuffix[i]; out[3 * steps + 2] = plan.relocatePrefix[i]; }
    for (const auto& sw : plan.swaps) { out[3 * steps] = 1; out[3 * steps + 1] = sw.first; out[3 * steps + 2] = sw.second; steps++; }
    return steps;
}
// gate queue: the steps flushGates() takes for a queue of one-target gates (layout.hpp planFlush), given the layout and the
// last-use stamps. logicalGates: target / ctrlMask name LOGICAL qubits. Outputs: steps = numSteps x {kind (0 = run of gates,
// 1 = relocation), count (gates | pairs), offset (first gate | first pair)}; pairs = {suffix bit, rank bit} per relocation pair;
// physTarget / physCtrlMask = every gate as the library gets it (index bits at the time it runs); whereOut = the layout afterwards.
// Buffers: steps 3 * (2 * numGates + 1), pairs 8 * (numGates + 1), physTarget / physCtrlMask numGates, whereOut numBits.
static unsigned writeFlushPlan(NatArray where, unsigned L, const std::vector<unsigned long long>& lastUse, const dfsa_gate1* gates, unsigned numGates,
                               unsigned* steps, unsigned* pairs, unsigned* physTarget, unsigned long long* physCtrlMask, unsigned* whereOut) {
    const std::vector<dfsa_detail::FlushStep> plan = dfsa_detail::planFlush(where, L, lastUse, gates, numGates);
    unsigned numPairs = 0, n = 0;
    for (const dfsa_detail::FlushStep& st : plan) {
        if (st.relocation) {
            steps[3 * n] = 1; steps[3 * n + 1] = unsigned(st.landing.size()); steps[3 * n + 2] = numPairs;
            for (std::size_t p = 0; p < st.landing.size(); p++, numPairs++) {
                pairs[2 * numPairs] = st.landing[p]; pairs[2 * numPairs + 1] = st.prefix[p];
                dfsa_detail::relabel(where, st.landing[p], st.prefix[p]);
            }
        } else {
            steps[3 * n] = 0; steps[3 * n + 1] = unsigned(st.count); steps[3 * n + 2] = unsigned(st.first);
            for (std::size_t g = st.first; g < st.first + st.count; g++) {
                const dfsa_gate1 phys = dfsa_detail::physicalGate(gates[g], where);
                physTarget[g] = phys.target; physCtrlMask[g] = phys.ctrlMask;
            }
        }
        n++;
    }
    for (std::size_t q = 0; q < where.size(); q++) whereOut[q] = where[q];
    return n;
}
unsigned dfsa_host_plan_flush(const unsigned* where, unsigned numBits, unsigned L, const unsigned long long* lastUse, const dfsa_gate1* logicalGates, unsigned numGates,
                              unsigned* steps, unsigned* pairs, unsigned* physTarget, unsigned long long* physCtrlMask, unsigned* whereOut) {
    return writeFlushPlan(NatArray(where, where + numBits), L, std::vector<unsigned long long>(lastUse, lastUse + numBits), logicalGates, numGates,
                          steps, pairs, physTarget, physCtrlMask, whereOut);
}
// ... for the gates a state has pending right now (what its next flushGates() will do); same buffers, sized with pendingGates
unsigned dfsa_host_state_planPendingFlush(void* h, unsigned* steps, unsigned* pairs, unsigned* physTarget, unsigned long long* physCtrlMask, unsigned* whereOut) {
    StateVector& s = sv(h);
    return writeFlushPlan(s.where, unsigned(s.logNumAmpsPerNode), s.lastUse, s.gateQueue.data(), unsigned(s.gateQueue.size()), steps, pairs, physTarget, physCtrlMask, whereOut);
}
// sortedTargets: the ket targets, ascending; reorderedOut has 2n entries, remainingOut 2N-2n
void dfsa_host_plan_partialTrace(unsigned N, unsigned L, const unsigned* sortedTargets, unsigned n, unsigned* reorderedOut, unsigned* remainingOut) {
    NatArray ext = toNats(sortedTargets, n);
    for (unsigned i = 0; i < n; i++) ext.push_back(sortedTargets[i] + N);
    NatArray re = getReorderedAllSuffixTargets(ext, L);
    NatArray rem = getNonTargetedQubitOrder(2 * N, ext, re);
    for (unsigned i = 0; i < 2 * n; i++) reorderedOut[i] = re[i];
    for (std::size_t i = 0; i < rem.size(); i++) remainingOut[i] = rem[i];
}

}  // extern "C"
