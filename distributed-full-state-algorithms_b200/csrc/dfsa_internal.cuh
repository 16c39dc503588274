// dfsa_internal.cuh -- shared by the translation units of libdfsa_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "dfsa_b200.h"

// ------------------------------------------------------------------------------------------------ state

struct dfsa_state {
    int      isDensity;
    unsigned numQubits;        // n (sv) or N (dm)
    int      rank, numNodes;
    unsigned logNumNodes;
    unsigned logNumAmps;       // per rank
    uint64_t numAmps;          // per rank
    double2* arr[2];           // DFSA_AMPS, DFSA_BUFFER (buffer == nullptr when numNodes == 1)
    int      allocId[2];       // slot in the IPC allocation registry (-1 if none); follows the arrays through swaps
    int      key;              // registry slot the shard was created with: stable name of this state across ranks
};

// ------------------------------------------------------------------------------------------------ context

enum class Transport { Single, Nccl, Ipc };

struct DfsaContext {
    bool         initialised = false;
    int          rank = 0, size = 1, device = 0;
    int          numSMs = 148;
    Transport    transport = Transport::Single;
    cudaStream_t compute = nullptr;   // every kernel
    cudaStream_t comm = nullptr;      // every transfer
    cudaEvent_t  evCompute = nullptr, evComm = nullptr;
    double2*     devScratch = nullptr;    // gate matrices, reduction partials
    size_t       devScratchBytes = 0;
    double*      hostPinned = nullptr;    // small pinned staging (reduction results)
};

DfsaContext& dfsaCtx();
extern uint64_t g_dfsaLaunches;                      // bumped at every kernel launch site
#define DFSA_COUNT_LAUNCH() (++g_dfsaLaunches)
void dfsaSetError(const char* fmt, ...);
int  dfsaEnsureDevice();                         // lazily create streams etc.; DFSA_ERR_CUDA if no device
int  dfsaScratch(size_t bytes, double2** out);   // grow-only device scratch
int  dfsaPoolDrain();                            // release the recycled small shards (collective when P > 1)

#define DFSA_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            dfsaSetError("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
            return DFSA_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define DFSA_REQUIRE(cond, msg)                                                                      \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            dfsaSetError("%s:%d: precondition failed: %s (%s)", __FILE__, __LINE__, #cond, msg);     \
            return DFSA_ERR_ARG;                                                                     \
        }                                                                                            \
    } while (0)

#define DFSA_TRY(call)                                                                               \
    do {                                                                                             \
        int r_ = (call);                                                                             \
        if (r_ != DFSA_OK) return r_;                                                                \
    } while (0)

#define DFSA_LAUNCH_CHECK() do { DFSA_COUNT_LAUNCH(); DFSA_CUDA(cudaGetLastError()); } while (0)

// range forms of the combine kernels (dfsa_kernels_sv.cu), used by the pipelined exchange in dfsa_comm.cu
int dfsaLaunchCombineRange(dfsa_state* s, uint64_t first, uint64_t num, double2 c0, double2 c1, bool partnerFirst);
int dfsaLaunchPauliCombineRange(dfsa_state* s, uint64_t first, uint64_t num, int pairRank, uint64_t maskXY, uint64_t maskYZ,
                                unsigned numY, double2 f, double2 h, bool exact);

// fused remote-load kernels (dfsa_kernels_sv.cu): buffer[j] = f0*amps[j] + f1*remote[j]  /  the Pauli form; the caller swaps arrays
int dfsaLaunchFusedCombine(dfsa_state* s, const double2* remote, double2 c0, double2 c1, bool partnerFirst);
int dfsaLaunchFusedPauliCombine(dfsa_state* s, const double2* remote, int pairRank, uint64_t maskXY, uint64_t maskYZ,
                                unsigned numY, double2 f, double2 h, bool exact);
int dfsaLaunchFusedSwap(dfsa_state* s, const double2* remote, unsigned qb, unsigned myBit);   // buffer = shard after swapping suffix qubit qb with this pair's prefix qubit
int dfsaLaunchFusedDepol1(dfsa_state* s, const double2* remote, unsigned qb, unsigned bit, double prob);    // prefix oneQubitDepolarising, buffer = result
int dfsaLaunchFusedDamping(dfsa_state* s, const double2* remote, unsigned qb, unsigned bit, double prob);   // prefix damping, buffer = result
int dfsaLaunchFusedCombineSub(dfsa_state* s, const double2* remote, const struct BitSpec& spec, uint64_t fixed, double2 c0, double2 c1, bool partnerFirst);
int dfsaCombineSub(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart, const double f0[2], const double f1[2], bool partnerFirst);   // buffer[j] = c0*amps[k(j)] + c1*remote[k(j)] on a control sub-cube
int dfsaLaunchFusedDepol2Pair(dfsa_state* s, const double2* remote, unsigned q0, unsigned q1, unsigned q2, unsigned bit, double prob, bool corrected);
int dfsaLaunchFusedDepol2Quad(dfsa_state* s, const double2* const* remote, unsigned q0, unsigned q1, unsigned bit0, unsigned bit1, double prob, bool corrected);
int dfsaLaunchRelocate(dfsa_state* s, const double2* const* peers, const unsigned* suffixPos, unsigned k, unsigned rho);   // buffer = shard after swapping k suffix with k prefix qubits
int dfsaPublishArrays(dfsa_state* s);            // tell the peers which registry slots are this state's amps / buffer now

// transport hooks implemented in dfsa_comm.cu
int dfsaRegisterAllocation(void* ptr, size_t bytes, int* idOut);   // collective when transport == Ipc
int dfsaUnregisterAllocation(int id);
int dfsaHostBarrier();                                             // inter-rank barrier without touching streams
// expecPauliString plumbing (dfsa_comm.cu): host-visible slot the reduction kernel publishes to, and the host-side collection
int dfsaExpecLocalOnly(bool on);
int dfsaExpecTarget(double** value, unsigned long long** flag, unsigned long long* seq, unsigned** ticket, int* global);
int dfsaExpecCollect(unsigned long long seq, double out[2]);
int dfsaAllreduceDoubles(double* v, int n, bool isMax);            // n <= 4 host values: rank-ordered sum, or max (NaN propagates)

// ------------------------------------------------------------------------------------------------ device helpers

// Sorted bit positions at which bits are inserted into a running index (bit_maths.hpp:37-58 restated for
// the device, 64-bit throughout -- the reference's Nat shifts truncate above 32 index bits).
struct BitSpec {
    uint32_t n;
    uint8_t  pos[DFSA_MAX_QUBITS];
};

__host__ __device__ __forceinline__ uint64_t insertZeroBit(uint64_t v, unsigned p) {
    uint64_t lowMask = (1ULL << p) - 1ULL;
    return ((v & ~lowMask) << 1) | (v & lowMask);
}

__host__ __device__ __forceinline__ uint64_t insertZeroBits(uint64_t v, const BitSpec& s) {
    for (uint32_t q = 0; q < s.n; q++) v = insertZeroBit(v, s.pos[q]);
    return v;
}

template <int N>
__device__ __forceinline__ uint64_t insertZeroBitsN(uint64_t v, const BitSpec& s) {
#pragma unroll
    for (int q = 0; q < N; q++) v = insertZeroBit(v, s.pos[q]);
    return v;
}

__device__ __forceinline__ unsigned parity64(uint64_t m) { return (unsigned)__popcll(m) & 1u; }

// complex arithmetic on double2 (x = re, y = im)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {   // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
// multiply by i^k without arithmetic: exact (bit-preserving up to the sign of zero)
__device__ __forceinline__ double2 mulPowI(double2 a, unsigned k) {
    switch (k & 3u) {
        case 0:  return a;
        case 1:  return make_double2(-a.y, a.x);
        case 2:  return make_double2(-a.x, -a.y);
        default: return make_double2(a.y, -a.x);
    }
}

struct Gate2 { double2 m00, m01, m10, m11; };

static inline double2 hostAmp(const double* p) { return make_double2(p[0], p[1]); }

// launch geometry: a grid that is a multiple of the SM count (persistent-style grid-stride kernels)
static inline unsigned dfsaGrid(uint64_t workItems, unsigned threads, unsigned itemsPerThread, unsigned blocksPerSM) {
    uint64_t perBlock = (uint64_t)threads * itemsPerThread;
    uint64_t needed = (workItems + perBlock - 1) / perBlock;
    uint64_t cap = (uint64_t)dfsaCtx().numSMs * blocksPerSM;
    if (needed < 1) needed = 1;
    return (unsigned)(needed < cap ? needed : cap);
}

// ------------------------------------------------------------------------------------------------ host-side helpers

#include <algorithm>
#include <vector>

// sorted, duplicate-free positions below `limit` -> BitSpec (+ their mask)
static inline int sortedSpec(const uint32_t* qubits, unsigned n, unsigned limit, BitSpec* spec, uint64_t* mask) {
    DFSA_REQUIRE(n <= DFSA_MAX_QUBITS, "too many qubits");
    std::vector<uint32_t> v(qubits, qubits + n);
    std::sort(v.begin(), v.end());
    uint64_t m = 0;
    for (unsigned q = 0; q < n; q++) {
        DFSA_REQUIRE(v[q] < limit, "qubit index is not a local (suffix) bit of this shard");
        DFSA_REQUIRE(q == 0 || v[q] != v[q - 1], "duplicate qubit");
        spec->pos[q] = (uint8_t)v[q];
        m |= 1ULL << v[q];
    }
    spec->n = n;
    if (mask) *mask = m;
    return DFSA_OK;
}

// spread the low bits of `value` over the sorted positions of `spec`
static inline uint64_t depositBits(uint64_t value, const BitSpec& spec) {
    uint64_t out = 0;
    for (uint32_t q = 0; q < spec.n; q++) out |= ((value >> q) & 1ULL) << spec.pos[q];
    return out;
}

// Pinned staging ring for small per-call host arguments (gate matrices, Pauli term tables): the caller fills the
// slot, enqueues its H2D copy on the compute stream and commits; a slot is reused only after its copy has run,
// so the host never has to block on the stream for argument lifetime.
constexpr size_t DFSA_STAGING_SLOT_BYTES = 128 * 1024;
int dfsaStagingAcquire(size_t bytes, void** hostPtr, int* slot);   // DFSA_ERR_UNSUPPORTED if bytes > slot size
int dfsaStagingCommit(int slot);
