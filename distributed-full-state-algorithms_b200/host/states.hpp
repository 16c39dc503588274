// states.hpp -- StateVector / DensityMatrix of the drop-in API, with the amplitudes resident in HBM.
// Same public fields and constructors as the reference's src/states.hpp (:16-25 fields, :32-48 ctor, :53-69
// DensityMatrix); `amps` and `buffer` are device-array views instead of std::vector. The test-only members the
// reference declares here and defines in tests/test_utilities.hpp (:419-524) are implemented with device copies.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <utility>
#include <vector>

#include "bit_maths.hpp"
#include "communication.hpp"
#include "misc.hpp"
#include "types.hpp"

#ifndef DFSA_AGREES_TOL
#define DFSA_AGREES_TOL 1E-5
#endif

class StateVector {
public:
    Nat rank = 0;
    Nat numNodes = 1;
    Nat logNumNodes = 0;

    Nat numQubits = 0;
    Index numAmpsPerNode = 0;
    Index logNumAmpsPerNode = 0;

    DeviceAmpArray amps;
    DeviceAmpArray buffer;

    dfsa_state* handle = nullptr;

    // lazy qubit layout (layout.hpp): where[q] = index bit currently holding logical qubit q; identity unless a gate left
    // its relocated targets in place. Anything that addresses amplitudes by index calls restoreLayout() first.
    NatArray where;
    std::vector<unsigned long long> lastUse;      // per logical qubit: when a one-target gate last touched it (layout.hpp: which qubit a swap-in evicts)
    unsigned long long useClock = 0;
    void touch(Nat logicalQubit) { lastUse[logicalQubit] = ++useClock; }
    bool layoutIsIdentity() const;
    NatArray physical(const NatArray& logical) const;
    void noteSwapped(Nat posA, Nat posB);
    Nat numDisplacedAcrossShardBoundary() const;
    void restoreLayout();

    // Deferred one-target gates (oneTargGate / manyCtrlOneTargGate, in call order; `target` and `ctrlMask` name LOGICAL qubits):
    // launched together by flushGates() so that consecutive gates share passes over HBM (dfsa_k_gateSequence) -- before any
    // other kind of operation on this state, before anything reads it, and at comm_synch(). flushGates() (layout.hpp) translates
    // them through the layout and, with the lazy layout on, brings the rank-bit qubits they target into the shard with as few
    // relocation steps as the queue allows (planFlush). DFSA_FUSE_GATES=0: never deferred.
    std::vector<dfsa_gate1> gateQueue;
    void flushGates();
    static bool& gateFusionEnabled() {
        static bool on = [] { const char* e = std::getenv("DFSA_FUSE_GATES"); return !(e && std::atoi(e) == 0); }();
        return on;
    }

    static std::vector<StateVector*>& registry() { static std::vector<StateVector*> live; return live; }      // every live state of this process
    static void flushAllStates() { for (StateVector* s : registry()) if (s->handle) s->flushGates(); }

    explicit StateVector(Nat numQubits) { create(false, numQubits); }
    virtual ~StateVector() { release(); }
    StateVector(const StateVector&) = delete;
    StateVector& operator=(const StateVector&) = delete;
    StateVector(StateVector&& other) noexcept { adopt(other); }
    StateVector& operator=(StateVector&& other) noexcept { if (this != &other) { release(); adopt(other); } return *this; }

    // ---- host <-> device (the reference's test-utility members + plain setters)
    AmpArray getAllVecAmps() {                                   // whole state on every rank
        restoreLayout();
        AmpArray all(Index(numNodes) * numAmpsPerNode);
        DFSA_CHECK(dfsa_state_download_all(handle, reinterpret_cast<double*>(all.data())));
        return all;
    }
    void setAllVecAmps(const AmpArray& all) {                    // each rank keeps its slice of the global array
        assert(all.size() == Index(numNodes) * numAmpsPerNode);
        resetLayout();
        comm_synch();
        DFSA_CHECK(dfsa_state_upload_all(handle, reinterpret_cast<const double*>(all.data())));
    }
    AmpArray getLocalAmps() {
        restoreLayout();
        AmpArray local(numAmpsPerNode);
        DFSA_CHECK(dfsa_state_download(handle, DFSA_AMPS, 0, numAmpsPerNode, reinterpret_cast<double*>(local.data())));
        return local;
    }
    void setLocalAmps(const AmpArray& local) {
        assert(local.size() == numAmpsPerNode);
        resetLayout();
        DFSA_CHECK(dfsa_state_upload(handle, DFSA_AMPS, 0, numAmpsPerNode, reinterpret_cast<const double*>(local.data())));
    }
    // Box-Muller on rand(), every rank drawing every rank's amplitudes so the streams stay in lock-step
    // (same sequence as tests/test_utilities.hpp:114-122, 438-448 for an identical seed)
    void setRandomAmps() {
        comm_synch();
        AmpArray mine(numAmpsPerNode);
        for (Nat r = 0; r < numNodes; r++)
            for (Index j = 0; j < numAmpsPerNode; j++) {
                Real a = std::rand() / Real(RAND_MAX), b = std::rand() / Real(RAND_MAX);
                Real radius = std::sqrt(-2 * std::log(a)), angle = 2 * 3.14159265 * b;
                if (r == rank) mine[j] = Amp(radius * std::cos(angle), radius * std::sin(angle));
            }
        setLocalAmps(mine);
    }
    void setHashAmps(unsigned long long seed) { resetLayout(); DFSA_CHECK(dfsa_state_init_hash(handle, seed)); }
    void resetLayout() { gateQueue.clear(); std::iota(where.begin(), where.end(), Nat(0)); std::fill(lastUse.begin(), lastUse.end(), 0ULL); useClock = 0; }      // the contents are about to be overwritten
    void printAmps() {
        AmpArray all = getAllVecAmps();
        if (rank == 0)
            for (Index i = 0; i < all.size(); i++) std::printf("%llu: (%.17g, %.17g)\n", i, all[i].real(), all[i].imag());
        comm_synch();
    }
    // two-sided and NaN-aware (the reference's comparison is one-sided and NaN-blind). Default tolerance
    // as the reference's (1E-5 absolute); a test build may tighten it: -DDFSA_AGREES_TOL=1e-12 -DDFSA_AGREES_RELATIVE makes the
    // bound tol * max(1, max |ref component|), the bar of BASELINE north_star.
    bool agreesWith(const AmpArray& ref, Real tol = DFSA_AGREES_TOL) {
        AmpArray all = getAllVecAmps();
        if (ref.size() != all.size()) return false;
        Real bound = tol;
#ifdef DFSA_AGREES_RELATIVE
        Real scale = 1;
        for (const Amp& a : ref) scale = std::max(scale, std::max(std::abs(a.real()), std::abs(a.imag())));
        bound = tol * scale;
#endif
        for (Index i = 0; i < ref.size(); i++) {
            Amp dif = ref[i] - all[i];
            if (!(std::abs(dif.real()) <= bound) || !(std::abs(dif.imag()) <= bound)) {
                if (rank == 0) std::printf("disagreement of (%g) + i(%g) at %llu (bound %g)\n", dif.real(), dif.imag(), i, bound);
                return false;
            }
        }
        return true;
    }
    Real getNorm2() { double n = 0; flushGates(); DFSA_CHECK(dfsa_state_norm2(handle, &n)); return n; }
    // device-resident utilities: nothing is gathered to the host
    void setPlusAmps() { resetLayout(); DFSA_CHECK(dfsa_state_init_plus(handle)); }
    void copyAmpsFrom(StateVector& other) { other.restoreLayout(); resetLayout(); DFSA_CHECK(dfsa_state_copy(handle, other.handle)); }
    // max over all amplitudes and ranks of |delta re|, |delta im| (NaN if any NaN); numUnequal counts == mismatches
    Real getMaxDifference(StateVector& other, Index* numUnequal = nullptr) {
        double d = 0; uint64_t ne = 0;
        restoreLayout();
        other.restoreLayout();
        DFSA_CHECK(dfsa_state_compare(handle, other.handle, &d, &ne, nullptr));
        if (numUnequal) *numUnequal = ne;
        return d;
    }
    bool agreesWith(StateVector& other, Real tol = DFSA_AGREES_TOL) { Real d = getMaxDifference(other); return d <= tol; }

protected:
    StateVector() = default;
    void create(bool isDensity, Nat qubits) {
        DFSA_CHECK(dfsa_state_create(isDensity ? 1 : 0, qubits, &handle));
        rank = comm_getRank();
        numNodes = comm_getNumNodes();
        logNumNodes = logBase2(numNodes);
        numQubits = qubits;
        logNumAmpsPerNode = dfsa_state_log_num_amps_per_node(handle);
        numAmpsPerNode = dfsa_state_num_amps_per_node(handle);
        amps = DeviceAmpArray{handle, DFSA_AMPS};
        buffer = DeviceAmpArray{handle, DFSA_BUFFER};
        where.resize(isDensity ? 2 * qubits : qubits);
        lastUse.resize(where.size());
        resetLayout();
        registry().push_back(this);
        dfsa_detail::flushAllStatesHook() = &StateVector::flushAllStates;
    }
    void unregister() { auto& r = registry(); r.erase(std::remove(r.begin(), r.end(), this), r.end()); }
    void release() {
        unregister();
        gateQueue.clear();                                          // nobody can observe their effect any more
        if (handle) DFSA_CHECK(dfsa_state_destroy(handle));
        handle = nullptr;
    }
    void adopt(StateVector& o) {
        rank = o.rank; numNodes = o.numNodes; logNumNodes = o.logNumNodes; numQubits = o.numQubits;
        numAmpsPerNode = o.numAmpsPerNode; logNumAmpsPerNode = o.logNumAmpsPerNode;
        amps = o.amps; buffer = o.buffer; handle = o.handle; where = std::move(o.where); gateQueue = std::move(o.gateQueue);
        lastUse = std::move(o.lastUse); useClock = o.useClock;
        if (handle) registry().push_back(this);
        o.unregister();
        o.handle = nullptr; o.amps = DeviceAmpArray(); o.buffer = DeviceAmpArray();
    }
};

// Choi vector of an N-qubit density matrix: 2N index bits, flat = 2^N*col + row; ranks own column blocks.
class DensityMatrix : public StateVector {
public:
    explicit DensityMatrix(Nat numQubits) : StateVector() { create(true, numQubits); }
    DensityMatrix(DensityMatrix&& other) noexcept : StateVector(std::move(other)) {}
    DensityMatrix& operator=(DensityMatrix&& other) noexcept { StateVector::operator=(std::move(other)); return *this; }

    AmpMatrix getAllMatrAmps() {
        const Index dim = powerOf2(numQubits);
        AmpArray vec = getAllVecAmps();
        AmpMatrix matr = getZeroMatrix(dim);
        for (Index c = 0; c < dim; c++)
            for (Index r = 0; r < dim; r++) matr[r][c] = vec[dim * c + r];
        return matr;
    }
    void setAllMatrAmps(const AmpMatrix& matr) {
        const Index dim = powerOf2(numQubits);
        AmpArray vec(dim * dim);
        for (Index c = 0; c < dim; c++)
            for (Index r = 0; r < dim; r++) vec[dim * c + r] = matr[r][c];
        setAllVecAmps(vec);
    }
    using StateVector::agreesWith;
    bool agreesWith(const AmpMatrix& ref, Real tol = DFSA_AGREES_TOL) {
        const Index dim = powerOf2(numQubits);
        if (ref.size() != dim) return false;
        AmpArray vec(dim * dim);
        for (Index c = 0; c < dim; c++)
            for (Index r = 0; r < dim; r++) vec[dim * c + r] = ref[r][c];
        return StateVector::agreesWith(vec, tol);
    }
};

#include "layout.hpp"
