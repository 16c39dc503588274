// dfsa_kernels_fused.cu -- a SEQUENCE of one-target (optionally controlled) gates in as few passes over HBM as possible.
//
// The reference applies every gate in its own sweep over the shard (src/local_statevector.hpp:14-29, :32-51: 32*A bytes per
// gate), and so do the per-gate kernels of dfsa_kernels_sv.cu, which sit at ~0.94 of the HBM roofline -- a per-gate bound no
// per-gate kernel can beat. A run of consecutive gates, however, only has to cross HBM once: here a tile of 2^11 amplitudes
// (32 KiB) spanned by the low index bits 0..3 (256-byte contiguous runs) and up to seven further target bits is staged in
// shared memory, EVERY gate of the batch whose target lies in the tile is applied to it there, in the caller's order, and the
// tile goes back. Controls may sit anywhere: inside the tile they predicate pairs, outside (other suffix bits, rank bits)
// they switch a gate on or off for a whole tile. Per-gate arithmetic is exactly that of the per-gate kernel
// (out0 = m00 a0 + m01 a1, out1 = m10 a0 + m11 a1, same FMA nesting), so the result is BIT-IDENTICAL to applying the gates
// one by one (tests/test_gpu_fused_gates.py compares the two on the device with == on every double).
//
// Inside a tile the gates are applied in GROUPS of consecutive gates whose targets fit three bits: a thread pulls the 8
// amplitudes those bits span into registers, applies the group's gates to them, and puts them back -- one shared-memory round
// trip per group instead of per gate (the 32-qubit sweep of bench.py: 6 gates per group). Slots are XOR-swizzled
// (low three element bits ^ bits 3..5 ^ bits 6..8) so that the eight lanes of a quarter-warp hit eight different 16-byte bank
// groups whichever three bits the group pulls out.
#include <algorithm>
#include <vector>

#include <string.h>
#include "dfsa_stream_kernels.cuh"

namespace {

constexpr unsigned FT_BITS = 11, FT_AMPS = 1u << FT_BITS, FT_THREADS = 256, FT_PER_THREAD = FT_AMPS / FT_THREADS;
constexpr unsigned FT_LOW_BITS = 4;                 // index bits 0..3 are always in the tile: 256-byte runs
constexpr unsigned FT_MAX_GATES = 64;

struct FusedGate {                                  // 96 bytes, read with 16-byte shared-memory loads
    double2  m00, m01, m10, m11;
    uint64_t ctrlExt;                               // controls outside the tile, as a mask on the GLOBAL index (rank bits included)
    uint32_t ctrlThread;                            // controls inside the tile but outside the group's three register bits: a mask on the
                                                    // tile-local element index -- the same for all 8 amplitudes a thread holds
    uint32_t rposAndCtrlReg;                        // bits 0-1: which register bit is the target; bits 4-6: which register bits are controls
    uint32_t pad[4];
};
struct FusedGroup { uint32_t firstGate, numGates, r0, r1, r2, pad; };   // r0 < r1 < r2: tile-local bit positions pulled into registers

__device__ __forceinline__ unsigned swz(unsigned e) { return e ^ (((e >> 3) ^ (e >> 6)) & 7u); }
__device__ __forceinline__ void ldsAmp2(double2& out, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out.x), "=d"(out.y) : "r"(addr));
}

template <unsigned RPOS>
__device__ __forceinline__ void applyInRegisters(double2 (&v)[8], double2 m00, double2 m01, double2 m10, double2 m11, unsigned ctrlReg) {
#pragma unroll
    for (unsigned p = 0; p < 4; p++) {
        // pair p: the two other register bits take the values of p's bits; i0 has the target bit clear
        const unsigned lo = p & ((1u << RPOS) - 1u), hi = p >> RPOS;
        const unsigned i0 = lo | (hi << (RPOS + 1)), i1 = i0 | (1u << RPOS);
        if ((i1 & ctrlReg) != ctrlReg) continue;                          // a control on one of the other two register bits is 0 for this pair
        const double2 a0 = v[i0], a1 = v[i1];
        v[i0] = cfma(m01, a1, cmul(m00, a0));
        v[i1] = cfma(m11, a1, cmul(m10, a0));
    }
}

// Shared memory per block: one tile slab and the pass's gate / group descriptors (read as warp-wide broadcasts: the first version
// fetched them from global memory per gate and per thread and spent a third of its stall samples waiting for those loads --
// profiles/r02_ncu_fused.txt). 64 registers and 39.5 KiB per block: FOUR blocks per SM, so one block's load / store / barrier
// phases are covered by the arithmetic of the other three (a second slab per block with prefetch was measured: no gain, and it
// limits the SM to three blocks).
constexpr unsigned FT_SLAB_BYTES = FT_AMPS * sizeof(double2);
constexpr unsigned FT_DESC_BYTES = FT_MAX_GATES * (sizeof(FusedGate) + sizeof(FusedGroup));
constexpr unsigned FT_SMEM_BYTES = FT_SLAB_BYTES + FT_DESC_BYTES;
constexpr unsigned FT_BLOCKS_PER_SM = 4;

__global__ void __launch_bounds__(FT_THREADS, FT_BLOCKS_PER_SM)
fusedGateTileKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec, const FusedGate* __restrict__ gatesGlobal, const FusedGroup* __restrict__ groupsGlobal,
                    unsigned numGates, unsigned numGroups, uint64_t rankShift) {
    extern __shared__ __align__(16) unsigned char fusedSmem[];
    double2* tile = reinterpret_cast<double2*>(fusedSmem);
    FusedGate* gates = reinterpret_cast<FusedGate*>(fusedSmem + FT_SLAB_BYTES);
    FusedGroup* groups = reinterpret_cast<FusedGroup*>(fusedSmem + FT_SLAB_BYTES + FT_MAX_GATES * sizeof(FusedGate));
    for (unsigned i = threadIdx.x; i < numGates * (sizeof(FusedGate) / 16); i += FT_THREADS)
        reinterpret_cast<uint4*>(gates)[i] = reinterpret_cast<const uint4*>(gatesGlobal)[i];
    for (unsigned i = threadIdx.x; i < numGroups * (sizeof(FusedGroup) / 8); i += FT_THREADS)
        reinterpret_cast<uint2*>(groups)[i] = reinterpret_cast<const uint2*>(groupsGlobal)[i];
    // element e = tid + 256 i: its offset inside a tile's span of the shard = (thread part) | (i part: tile bits 8..10)
    uint64_t offTid = 0;
#pragma unroll
    for (unsigned b = 0; b < 8; b++) offTid |= (uint64_t)((threadIdx.x >> b) & 1u) << tileSpec.pos[b];
    const uint64_t hi0 = 1ULL << tileSpec.pos[8], hi1 = 1ULL << tileSpec.pos[9], hi2 = 1ULL << tileSpec.pos[10];
    auto off = [&](unsigned i) { return offTid | ((i & 1u) ? hi0 : 0ULL) | ((i & 2u) ? hi1 : 0ULL) | ((i & 4u) ? hi2 : 0ULL); };
    for (uint64_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
        const uint64_t base = insertZeroBitsN<FT_BITS>(t, tileSpec);
#pragma unroll
        for (unsigned i = 0; i < FT_PER_THREAD; i++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[swz(threadIdx.x + FT_THREADS * i)]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(amps + (base | off(i))) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const uint64_t gidx = rankShift | base;
        for (unsigned gi = 0; gi < numGroups; gi++) {
            const FusedGroup grp = groups[gi];
            // this thread's octet: the 8 tile bits that are not pulled out come from the thread index
            unsigned eb = threadIdx.x;
            eb = ((eb >> grp.r0) << (grp.r0 + 1)) | (eb & ((1u << grp.r0) - 1u));
            eb = ((eb >> grp.r1) << (grp.r1 + 1)) | (eb & ((1u << grp.r1) - 1u));
            eb = ((eb >> grp.r2) << (grp.r2 + 1)) | (eb & ((1u << grp.r2) - 1u));
            const unsigned b0 = 1u << grp.r0, b1 = 1u << grp.r1, b2 = 1u << grp.r2;
            double2 v[8];
#pragma unroll
            for (unsigned i = 0; i < 8; i++) v[i] = tile[swz(eb | ((i & 1u) ? b0 : 0u) | ((i & 2u) ? b1 : 0u) | ((i & 4u) ? b2 : 0u))];
            unsigned gAddr = (unsigned)__cvta_generic_to_shared(&gates[grp.firstGate]);
            for (unsigned k = 0; k < grp.numGates; k++, gAddr += (unsigned)sizeof(FusedGate)) {
                // the descriptor comes as five 16-byte broadcast loads straight into vector registers (left to the compiler it went
                // through uniform registers: a dozen R2UR + as many moves per gate, profiles/r02_ncu_fused.txt)
                unsigned long long ctrlExt, second;                       // second = ctrlThread | rposAndCtrlReg << 32
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(ctrlExt), "=l"(second) : "r"(gAddr + 64u));
                double2 m00, m01, m10, m11;
                ldsAmp2(m00, gAddr); ldsAmp2(m01, gAddr + 16u); ldsAmp2(m10, gAddr + 32u); ldsAmp2(m11, gAddr + 48u);
                const unsigned ctrlThread = (unsigned)second, packed = (unsigned)(second >> 32);
                if ((gidx & ctrlExt) != ctrlExt) continue;                // a control outside the tile is 0 for this whole tile
                if ((eb & ctrlThread) != ctrlThread) continue;            // ... or for everything this thread holds
                const unsigned rpos = packed & 3u, ctrlReg = (packed >> 4) & 7u;
                if (rpos == 0) applyInRegisters<0>(v, m00, m01, m10, m11, ctrlReg);
                else if (rpos == 1) applyInRegisters<1>(v, m00, m01, m10, m11, ctrlReg);
                else applyInRegisters<2>(v, m00, m01, m10, m11, ctrlReg);
            }
#pragma unroll
            for (unsigned i = 0; i < 8; i++) tile[swz(eb | ((i & 1u) ? b0 : 0u) | ((i & 2u) ? b1 : 0u) | ((i & 4u) ? b2 : 0u))] = v[i];
            __syncthreads();
        }
#pragma unroll
        for (unsigned i = 0; i < FT_PER_THREAD; i++) amps[base | off(i)] = tile[swz(threadIdx.x + FT_THREADS * i)];
        __syncthreads();                                                  // the next tile's copies overwrite the slab
    }
}

// ---- planning (host, pure): which gates share a pass, which bits the pass's tile spans, how its gates group into octets
struct PlannedBatch {
    unsigned first = 0, count = 0;                  // gates [first, first + count) of the sequence
    bool tiled = false;                             // false: a single gate, goes to the per-gate stream kernel
    std::vector<uint32_t> tileBits;                 // ascending index bits of the tile
    std::vector<FusedGroup> groups;                 // r0..r2 as TILE-LOCAL positions
    std::vector<uint32_t> rpos;                     // per gate: which register bit is its target
};

void planGateSequence(const dfsa_gate1* seq, unsigned n, unsigned L, std::vector<PlannedBatch>* out) {
    out->clear();
    unsigned i = 0;
    while (i < n) {
        PlannedBatch b;
        b.first = i;
        uint64_t S = (1ULL << FT_LOW_BITS) - 1ULL;
        unsigned j = i;
        if (L >= FT_BITS)
            while (j < n && j - i < FT_MAX_GATES && __builtin_popcountll(S | (1ULL << seq[j].target)) <= (int)FT_BITS) { S |= 1ULL << seq[j].target; j++; }
        if (j - i <= 1) { b.count = 1; b.tiled = false; out->push_back(b); i++; continue; }
        b.count = j - i;
        b.tiled = true;
        for (unsigned q = 0; q < L && __builtin_popcountll(S) < (int)FT_BITS; q++) S |= 1ULL << q;      // fill up with the lowest free bits
        for (unsigned q = 0; q < L; q++) if ((S >> q) & 1ULL) b.tileBits.push_back(q);
        auto localPos = [&](unsigned q) { return (unsigned)(std::find(b.tileBits.begin(), b.tileBits.end(), q) - b.tileBits.begin()); };
        // groups: maximal runs of consecutive gates whose targets span at most three bits; padded to three with the lowest other tile bits
        unsigned k = i;
        while (k < j) {
            std::vector<unsigned> R;
            unsigned e = k;
            while (e < j) {
                const unsigned p = localPos(seq[e].target);
                if (std::find(R.begin(), R.end(), p) == R.end()) { if (R.size() == 3) break; R.push_back(p); }
                e++;
            }
            for (unsigned p = 0; R.size() < 3; p++) if (std::find(R.begin(), R.end(), p) == R.end()) R.push_back(p);
            std::sort(R.begin(), R.end());
            FusedGroup g{k - i, e - k, R[0], R[1], R[2], 0};
            b.groups.push_back(g);
            for (unsigned x = k; x < e; x++) b.rpos.push_back((unsigned)(std::find(R.begin(), R.end(), localPos(seq[x].target)) - R.begin()));
            k = e;
        }
        out->push_back(b);
        i = j;
    }
}

bool fusionEnabledInLibrary() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("DFSA_FUSE_GATES"); on = (e && atoi(e) == 0) ? 0 : 1; }
    return on == 1;
}

}  // namespace

// Host-only view of the plan (no device needed): for every gate the pass (batch) it runs in, whether that pass is a tile pass,
// and the group inside it; tileBitsOut receives 11 index bits per batch. tests/test_fused_plan.py checks that every gate's target
// lies in its pass's tile, that groups span <= 3 bits and that order is preserved.
extern "C" int dfsa_plan_gateSequence(const dfsa_gate1* gates, unsigned numGates, unsigned logNumAmps, uint32_t* batchOfGate, uint32_t* groupOfGate,
                                      uint32_t* batchIsTiled, uint32_t* tileBitsOut, uint32_t* groupBitsOut, unsigned* numBatches) {
    DFSA_REQUIRE(gates && batchOfGate && groupOfGate && batchIsTiled && tileBitsOut && groupBitsOut && numBatches, "null argument");
    for (unsigned g = 0; g < numGates; g++) DFSA_REQUIRE(gates[g].target < logNumAmps && !((gates[g].ctrlMask >> gates[g].target) & 1ULL), "bad gate");
    std::vector<PlannedBatch> plan;
    planGateSequence(gates, numGates, logNumAmps, &plan);
    for (size_t b = 0; b < plan.size(); b++) {
        batchIsTiled[b] = plan[b].tiled ? 1u : 0u;
        for (unsigned q = 0; q < FT_BITS; q++) tileBitsOut[b * FT_BITS + q] = plan[b].tiled ? plan[b].tileBits[q] : 0u;
        for (unsigned x = 0; x < plan[b].count; x++) { batchOfGate[plan[b].first + x] = (uint32_t)b; groupOfGate[plan[b].first + x] = 0; }
        for (size_t gi = 0; gi < plan[b].groups.size(); gi++) {
            const FusedGroup& g = plan[b].groups[gi];
            for (unsigned x = 0; x < g.numGates; x++) {
                const unsigned idx = plan[b].first + g.firstGate + x;
                groupOfGate[idx] = (uint32_t)gi;
                groupBitsOut[3 * idx] = plan[b].tileBits[g.r0]; groupBitsOut[3 * idx + 1] = plan[b].tileBits[g.r1]; groupBitsOut[3 * idx + 2] = plan[b].tileBits[g.r2];
            }
        }
    }
    *numBatches = (unsigned)plan.size();
    return DFSA_OK;
}

// K1/K2 for a whole run of gates (local_statevector.hpp:14-51 applied numGates times): every target a suffix bit, controls
// anywhere (a control on a rank bit that this rank fails drops the gate). Same results, bit for bit, as numGates calls of
// dfsa_k_ctrlOneTarg; consecutive gates share passes over HBM.
extern "C" int dfsa_k_gateSequence(dfsa_state* s, const dfsa_gate1* gates, unsigned numGates) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && (gates || numGates == 0), "null argument");
    const unsigned L = s->logNumAmps;
    const uint64_t localMask = (1ULL << L) - 1ULL, rankShift = (uint64_t)s->rank << L;
    std::vector<dfsa_gate1> live;                   // gates this rank takes part in (all its rank-bit controls are 1)
    live.reserve(numGates);
    for (unsigned g = 0; g < numGates; g++) {
        DFSA_REQUIRE(gates[g].target < L, "targets must be suffix bits");
        DFSA_REQUIRE(!((gates[g].ctrlMask >> gates[g].target) & 1ULL), "a gate cannot be controlled on its own target");
        const uint64_t pre = gates[g].ctrlMask & ~localMask;
        if ((rankShift & pre) == pre) live.push_back(gates[g]);
    }
    auto perGate = [&](const dfsa_gate1& g) {
        uint32_t ctrls[64]; unsigned nc = 0;
        for (unsigned q = 0; q < L; q++) if ((g.ctrlMask >> q) & 1ULL) ctrls[nc++] = q;
        return dfsa_k_ctrlOneTarg(s, ctrls, nc, g.target, g.matrix);
    };
    if (!fusionEnabledInLibrary()) { for (const dfsa_gate1& g : live) DFSA_TRY(perGate(g)); return DFSA_OK; }
    std::vector<PlannedBatch> plan;
    planGateSequence(live.data(), (unsigned)live.size(), L, &plan);
    DfsaContext& ctx = dfsaCtx();
    for (const PlannedBatch& b : plan) {
        if (!b.tiled) { DFSA_TRY(perGate(live[b.first])); continue; }
        BitSpec tileSpec; tileSpec.n = FT_BITS;
        uint64_t tileMask = 0;
        for (unsigned q = 0; q < FT_BITS; q++) { tileSpec.pos[q] = (uint8_t)b.tileBits[q]; tileMask |= 1ULL << b.tileBits[q]; }
        const size_t gateBytes = sizeof(FusedGate) * b.count, groupBytes = sizeof(FusedGroup) * b.groups.size();
        void* stage; int slot;
        DFSA_TRY(dfsaStagingAcquire(gateBytes + groupBytes, &stage, &slot));
        FusedGate* hg = (FusedGate*)stage;
        for (unsigned x = 0; x < b.count; x++) {
            const dfsa_gate1& src = live[b.first + x];
            FusedGate& d = hg[x];
            d.m00 = hostAmp(src.matrix); d.m01 = hostAmp(src.matrix + 2); d.m10 = hostAmp(src.matrix + 4); d.m11 = hostAmp(src.matrix + 6);
            const uint64_t local = src.ctrlMask & localMask;
            d.ctrlExt = (local & ~tileMask) | (src.ctrlMask & ~localMask);     // satisfied rank-bit controls test true against rankShift
            unsigned ctrlTile = 0;
            for (unsigned q = 0; q < FT_BITS; q++) if ((local >> b.tileBits[q]) & 1ULL) ctrlTile |= 1u << q;
            // the group this gate belongs to decides which tile bits live in registers
            const FusedGroup* grp = nullptr;
            for (const FusedGroup& gg : b.groups) if (x >= gg.firstGate && x < gg.firstGate + gg.numGates) grp = &gg;
            const unsigned rbit[3] = {grp->r0, grp->r1, grp->r2};
            unsigned ctrlReg = 0;
            for (unsigned r = 0; r < 3; r++) if ((ctrlTile >> rbit[r]) & 1u) { ctrlReg |= 1u << r; ctrlTile &= ~(1u << rbit[r]); }
            d.ctrlThread = ctrlTile;
            d.rposAndCtrlReg = b.rpos[x] | (ctrlReg << 4);
            memset(d.pad, 0, sizeof(d.pad));
        }
        memcpy((char*)stage + gateBytes, b.groups.data(), groupBytes);
        double2* dev;
        DFSA_TRY(dfsaScratch(gateBytes + groupBytes, &dev));
        DFSA_CUDA(cudaMemcpyAsync(dev, stage, gateBytes + groupBytes, cudaMemcpyHostToDevice, ctx.compute));
        DFSA_TRY(dfsaStagingCommit(slot));
        const uint64_t numTiles = s->numAmps >> FT_BITS;
        const unsigned grid = (unsigned)std::min<uint64_t>(numTiles, (uint64_t)ctx.numSMs * FT_BLOCKS_PER_SM);
        static bool configured = false;
        if (!configured) {
            DFSA_CUDA(cudaFuncSetAttribute(fusedGateTileKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FT_SMEM_BYTES));
            configured = true;
        }
        fusedGateTileKernel<<<grid, FT_THREADS, FT_SMEM_BYTES, ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, (const FusedGate*)dev,
                                                                              (const FusedGroup*)((const char*)dev + gateBytes), b.count, (unsigned)b.groups.size(), rankShift);
        DFSA_LAUNCH_CHECK();
    }
    return DFSA_OK;
}
