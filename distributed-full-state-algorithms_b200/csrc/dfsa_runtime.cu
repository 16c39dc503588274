// dfsa_runtime.cu -- context, error reporting and state storage of libdfsa_b200.so.
// Replaces the std::vector storage of the reference's StateVector/DensityMatrix (src/states.hpp:13-69):
// each rank's amplitude shard and its equal-size exchange buffer are plain cudaMalloc regions in HBM.
#include <stdarg.h>
#include <math.h>
#include <string.h>
#include <vector>

#include "dfsa_stream_kernels.cuh"

static thread_local char g_err[1024] = "";
static DfsaContext g_ctx;

uint64_t g_dfsaLaunches = 0;

DfsaContext& dfsaCtx() { return g_ctx; }

void dfsaSetError(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("DFSA_VERBOSE")) fprintf(stderr, "[dfsa rank %d] %s\n", g_ctx.rank, g_err);
}

extern "C" const char* dfsa_last_error(void) { return g_err; }
extern "C" const char* dfsa_version(void) { return "dfsa_b200 0.1 (sm_100a)"; }

int dfsaEnsureDevice() {
    DfsaContext& c = g_ctx;
    if (c.compute) return DFSA_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        dfsaSetError("no CUDA device available (%s); libdfsa_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return DFSA_ERR_CUDA;
    }
    if (!c.initialised) c.device = 0;   // single-rank use without dfsa_comm_init
    DFSA_CUDA(cudaSetDevice(c.device % count));
    cudaDeviceProp prop;
    DFSA_CUDA(cudaGetDeviceProperties(&prop, c.device % count));
    c.numSMs = prop.multiProcessorCount;
    DFSA_CUDA(cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking));
    DFSA_CUDA(cudaStreamCreateWithFlags(&c.comm, cudaStreamNonBlocking));
    DFSA_CUDA(cudaEventCreateWithFlags(&c.evCompute, cudaEventDisableTiming));
    DFSA_CUDA(cudaEventCreateWithFlags(&c.evComm, cudaEventDisableTiming));
    DFSA_CUDA(cudaHostAlloc((void**)&c.hostPinned, 4096, cudaHostAllocDefault));
    return DFSA_OK;
}

int dfsaScratch(size_t bytes, double2** out) {
    DfsaContext& c = g_ctx;
    if (bytes > c.devScratchBytes) {
        // the previous scratch may still be read by an enqueued kernel
        DFSA_CUDA(cudaStreamSynchronize(c.compute));
        if (c.devScratch) DFSA_CUDA(cudaFree(c.devScratch));
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        DFSA_CUDA(cudaMalloc((void**)&c.devScratch, want));
        c.devScratchBytes = want;
    }
    *out = c.devScratch;
    return DFSA_OK;
}

namespace {
constexpr int kStagingSlots = 8;
struct StagingRing { char* host = nullptr; cudaEvent_t done[kStagingSlots]; bool used[kStagingSlots] = {}; int next = 0; } g_staging;
}

int dfsaStagingAcquire(size_t bytes, void** hostPtr, int* slot) {
    if (bytes > DFSA_STAGING_SLOT_BYTES) return DFSA_ERR_UNSUPPORTED;
    if (!g_staging.host) {
        DFSA_CUDA(cudaHostAlloc((void**)&g_staging.host, DFSA_STAGING_SLOT_BYTES * kStagingSlots, cudaHostAllocDefault));
        for (int i = 0; i < kStagingSlots; i++) DFSA_CUDA(cudaEventCreateWithFlags(&g_staging.done[i], cudaEventDisableTiming));
    }
    int i = g_staging.next;
    g_staging.next = (i + 1) % kStagingSlots;
    if (g_staging.used[i]) DFSA_CUDA(cudaEventSynchronize(g_staging.done[i]));
    *hostPtr = g_staging.host + (size_t)i * DFSA_STAGING_SLOT_BYTES;
    *slot = i;
    return DFSA_OK;
}

int dfsaStagingCommit(int slot) {
    DFSA_CUDA(cudaEventRecord(g_staging.done[slot], g_ctx.compute));
    g_staging.used[slot] = true;
    return DFSA_OK;
}

extern "C" void* dfsa_stream_compute(void) { return dfsaEnsureDevice() == DFSA_OK ? (void*)g_ctx.compute : nullptr; }

extern "C" int dfsa_device_sync(void) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.comm));
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.compute));
    return DFSA_OK;
}

// ------------------------------------------------------------------------------------------------ measurement helpers

extern "C" int dfsa_event_create(void** event) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(event, "null argument");
    cudaEvent_t e;
    DFSA_CUDA(cudaEventCreate(&e));
    *event = (void*)e;
    return DFSA_OK;
}
extern "C" int dfsa_event_record(void* event) {
    DFSA_REQUIRE(event, "null event");
    DFSA_CUDA(cudaEventRecord((cudaEvent_t)event, g_ctx.compute));
    return DFSA_OK;
}
extern "C" int dfsa_event_elapsed_ms(void* start, void* stop, double* ms) {
    DFSA_REQUIRE(start && stop && ms, "null argument");
    DFSA_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    float f = 0.f;
    DFSA_CUDA(cudaEventElapsedTime(&f, (cudaEvent_t)start, (cudaEvent_t)stop));
    *ms = f;
    return DFSA_OK;
}
extern "C" int dfsa_event_destroy(void* event) {
    if (event) DFSA_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return DFSA_OK;
}
extern "C" uint64_t dfsa_launch_count(void) { return g_dfsaLaunches; }

extern "C" int dfsa_host_alloc_pinned(uint64_t bytes, void** out) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(out && bytes > 0, "bad argument");
    DFSA_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return DFSA_OK;
}
extern "C" int dfsa_host_free_pinned(void* ptr) {
    if (ptr) DFSA_CUDA(cudaFreeHost(ptr));
    return DFSA_OK;
}

// ------------------------------------------------------------------------------------------------ states

// Small shards are recycled instead of going back to the driver: partialTrace returns a NEW DensityMatrix per call
// (distributed_densitymatrix.hpp:347), and a cudaMalloc + IPC export + peer mapping per call costs far more than the
// trace itself (round 1, 8 GPUs: 60 ms around a 0.1 ms kernel). Pooled arrays keep their registry slots and peer
// mappings. Every rank frees and creates the same sizes in the same order (SPMD), so the pools stay in step.
namespace {
struct PooledShard { size_t bytes; int numArrays; double2* arr[2]; int allocId[2]; };
std::vector<PooledShard> g_pool;
constexpr size_t kPoolMaxShardBytes = 256u << 20;     // per array
constexpr size_t kPoolMaxEntries = 8;
}

int dfsaPoolDrain() {
    for (PooledShard& e : g_pool) {
        if (g_ctx.size > 1)
            for (int w = 0; w < 2; w++) if (e.allocId[w] >= 0) DFSA_TRY(dfsaUnregisterAllocation(e.allocId[w]));
        for (int w = 0; w < 2; w++) if (e.arr[w]) DFSA_CUDA(cudaFree(e.arr[w]));
    }
    g_pool.clear();
    return DFSA_OK;
}

extern "C" int dfsa_state_create(int isDensity, unsigned numQubits, dfsa_state** out) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(out, "null out pointer");
    DfsaContext& c = g_ctx;
    unsigned k = 0;
    while ((1 << k) < c.size) k++;
    DFSA_REQUIRE((1 << k) == c.size, "number of ranks must be a power of 2 (README.md:134)");
    // 0 qubits (a single amplitude) is legal: partialTrace may trace out every qubit at 1 rank
    DFSA_REQUIRE(numQubits < 63 && (numQubits >= 31 || (1ULL << numQubits) >= (uint64_t)c.size),
                 "need 2^numQubits >= number of ranks (states.hpp:35)");
    unsigned bits = isDensity ? 2 * numQubits : numQubits;
    DFSA_REQUIRE(bits < 63 && bits >= k, "state too large / too small for this many ranks");
    dfsa_state* s = new dfsa_state();
    s->isDensity = isDensity ? 1 : 0;
    s->numQubits = numQubits;
    s->rank = c.rank;
    s->numNodes = c.size;
    s->logNumNodes = k;
    s->logNumAmps = bits - k;
    s->numAmps = 1ULL << s->logNumAmps;
    s->arr[0] = s->arr[1] = nullptr;
    s->allocId[0] = s->allocId[1] = -1;
    s->key = -1;
    size_t bytes = s->numAmps * sizeof(double2);
    int numArrays = (c.size > 1) ? 2 : 1;          // the exchange buffer is only ever touched when P > 1
    bool recycled = false;
    for (size_t i = 0; i < g_pool.size(); i++) {
        if (g_pool[i].bytes != bytes || g_pool[i].numArrays != numArrays) continue;
        for (int w = 0; w < 2; w++) { s->arr[w] = g_pool[i].arr[w]; s->allocId[w] = g_pool[i].allocId[w]; }
        g_pool.erase(g_pool.begin() + i);
        recycled = true;
        break;
    }
    for (int w = 0; w < numArrays; w++) {
        if (!recycled) {
            cudaError_t e = cudaMalloc((void**)&s->arr[w], bytes);
            if (e != cudaSuccess) {
                dfsaSetError("cudaMalloc of %zu bytes for the %s failed: %s", bytes, w ? "exchange buffer" : "amplitude shard", cudaGetErrorString(e));
                if (s->arr[0]) cudaFree(s->arr[0]);
                delete s;
                return DFSA_ERR_CUDA;
            }
        }
        DFSA_CUDA(cudaMemsetAsync(s->arr[w], 0, bytes, c.compute));
    }
    if (c.size > 1) {
        if (!recycled)
            for (int w = 0; w < 2; w++) DFSA_TRY(dfsaRegisterAllocation(s->arr[w], bytes, &s->allocId[w]));
        s->key = s->allocId[0];
        DFSA_TRY(dfsaPublishArrays(s));
    }
    *out = s;
    return DFSA_OK;
}

extern "C" int dfsa_state_destroy(dfsa_state* s) {
    if (!s) return DFSA_OK;
    DFSA_TRY(dfsa_device_sync());
    const size_t bytes = s->numAmps * sizeof(double2);
    if (bytes <= kPoolMaxShardBytes && g_pool.size() < kPoolMaxEntries) {
        PooledShard e{bytes, s->arr[1] ? 2 : 1, {s->arr[0], s->arr[1]}, {s->allocId[0], s->allocId[1]}};
        // amps and buffer may have traded places a different number of times on different ranks (a control-gated rank
        // skips the exchange): order the pair by registry slot, which IS the same everywhere, so that the next owner's
        // key (= allocId[0]) agrees across ranks
        if (e.arr[1] && e.allocId[1] < e.allocId[0]) { std::swap(e.arr[0], e.arr[1]); std::swap(e.allocId[0], e.allocId[1]); }
        g_pool.push_back(e);
        delete s;
        return DFSA_OK;
    }
    if (g_ctx.size > 1)
        for (int w = 0; w < 2; w++) if (s->allocId[w] >= 0) DFSA_TRY(dfsaUnregisterAllocation(s->allocId[w]));
    for (int w = 0; w < 2; w++) if (s->arr[w]) DFSA_CUDA(cudaFree(s->arr[w]));
    delete s;
    return DFSA_OK;
}

extern "C" double* dfsa_state_ptr(dfsa_state* s, int which) { return (s && (which == 0 || which == 1)) ? (double*)s->arr[which] : nullptr; }
extern "C" uint64_t dfsa_state_num_amps_per_node(const dfsa_state* s) { return s ? s->numAmps : 0; }
extern "C" unsigned dfsa_state_log_num_amps_per_node(const dfsa_state* s) { return s ? s->logNumAmps : 0; }
extern "C" unsigned dfsa_state_num_qubits(const dfsa_state* s) { return s ? s->numQubits : 0; }
extern "C" int dfsa_state_is_density(const dfsa_state* s) { return s ? s->isDensity : 0; }

extern "C" int dfsa_state_swap_arrays(dfsa_state* s) {
    DFSA_REQUIRE(s && s->arr[1], "no exchange buffer to swap with");
    std::swap(s->arr[0], s->arr[1]);
    std::swap(s->allocId[0], s->allocId[1]);
    return dfsaPublishArrays(s);
}

extern "C" int dfsa_state_upload(dfsa_state* s, int which, uint64_t first, uint64_t num, const double* host) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && host && (which == 0 || which == 1) && s->arr[which], "bad argument");
    DFSA_REQUIRE(first + num <= s->numAmps, "range exceeds the shard");
    DFSA_CUDA(cudaMemcpyAsync(s->arr[which] + first, host, num * sizeof(double2), cudaMemcpyHostToDevice, g_ctx.compute));
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.compute));
    return DFSA_OK;
}

extern "C" int dfsa_state_download(dfsa_state* s, int which, uint64_t first, uint64_t num, double* host) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && host && (which == 0 || which == 1) && s->arr[which], "bad argument");
    DFSA_REQUIRE(first + num <= s->numAmps, "range exceeds the shard");
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.comm));
    DFSA_CUDA(cudaMemcpyAsync(host, s->arr[which] + first, num * sizeof(double2), cudaMemcpyDeviceToHost, g_ctx.compute));
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.compute));
    return DFSA_OK;
}

extern "C" int dfsa_state_upload_all(dfsa_state* s, const double* hostAll) {
    DFSA_REQUIRE(s && hostAll, "null argument");
    return dfsa_state_upload(s, DFSA_AMPS, 0, s->numAmps, hostAll + 2 * (uint64_t)s->rank * s->numAmps);
}

extern "C" int dfsa_state_init_zero(dfsa_state* s) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    DFSA_CUDA(cudaMemsetAsync(s->arr[0], 0, s->numAmps * sizeof(double2), g_ctx.compute));
    return DFSA_OK;
}

// the synthetic state of SURVEY 8(d): amp(i) = hash(seed, global index i), reproducible on the host
// (oracle/dfsa_oracle.c: orc_init_hash, oracle/ref_driver.cpp: hashReal)
__device__ __forceinline__ double hashReal(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1ULL) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

extern "C" int dfsa_state_init_hash(dfsa_state* s, uint64_t seed) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    double2* amps = s->arr[0];
    uint64_t first = (uint64_t)s->rank << s->logNumAmps;
    auto ld = [=] __device__(uint64_t j) { return Amp1{make_double2(hashReal(seed, 2 * (first + j)), hashReal(seed, 2 * (first + j) + 1))}; };
    auto st = [=] __device__(uint64_t j, const Amp1& v) { amps[j] = v.a; };
    return launchStream<2, Amp1>(s->numAmps, ld, st);
}

extern "C" int dfsa_state_init_plus(dfsa_state* s) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    double2* amps = s->arr[0];
    // |+>^n: 2^(-n/2) everywhere; as a density matrix |+><+|^N: 2^(-N) everywhere (both exact powers of two for even n)
    const double v = s->isDensity ? ldexp(1.0, -(int)s->numQubits) : pow(2.0, -0.5 * s->numQubits);
    auto ld = [=] __device__(uint64_t) { return Amp1{make_double2(v, 0.0)}; };
    auto st = [=] __device__(uint64_t j, const Amp1& a) { amps[j] = a.a; };
    return launchStream<2, Amp1>(s->numAmps, ld, st);
}

extern "C" int dfsa_state_copy(dfsa_state* dst, const dfsa_state* src) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(dst && src && dst->numAmps == src->numAmps && dst->isDensity == src->isDensity, "states must have the same shape");
    DFSA_CUDA(cudaMemcpyAsync(dst->arr[0], src->arr[0], src->numAmps * sizeof(double2), cudaMemcpyDeviceToDevice, g_ctx.compute));
    return DFSA_OK;
}

// On-device comparator (the reference's agreesWith gathers both states to every rank's host, test_utilities.hpp:470-487):
// per block {max |delta component|, max |reference component|, number of unequal amplitudes, NaN seen}; the reference side is
// either a second shard or the hash state regenerated on the fly.
struct CmpPartial { double maxDiff, maxRef, unequal, nan; };

template <bool HASH>
__global__ void __launch_bounds__(256) compareKernel(const double2* __restrict__ a, const double2* __restrict__ b, uint64_t n, uint64_t seed,
                                                    uint64_t first, CmpPartial* partials) {
    double md = 0.0, mr = 0.0;
    unsigned long long ne = 0;
    unsigned nan = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double2 x = a[i];
        double2 y;
        if (HASH) y = make_double2(hashReal(seed, 2 * (first + i)), hashReal(seed, 2 * (first + i) + 1));
        else y = b[i];
        const double dx = fabs(x.x - y.x), dy = fabs(x.y - y.y);
        nan |= (dx != dx) | (dy != dy);
        md = fmax(md, fmax(dx, dy));
        mr = fmax(mr, fmax(fabs(y.x), fabs(y.y)));
        ne += (x.x != y.x) | (x.y != y.y);
    }
    __shared__ CmpPartial w[8];
    double ned = (double)ne, nand = (double)nan;
    for (int off = 16; off > 0; off >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, off));
        mr = fmax(mr, __shfl_xor_sync(0xffffffffu, mr, off));
        ned += __shfl_xor_sync(0xffffffffu, ned, off);
        nand += __shfl_xor_sync(0xffffffffu, nand, off);
    }
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = CmpPartial{md, mr, ned, nand};
    __syncthreads();
    if (threadIdx.x == 0) {
        CmpPartial t = w[0];
        for (int k = 1; k < 8; k++) { t.maxDiff = fmax(t.maxDiff, w[k].maxDiff); t.maxRef = fmax(t.maxRef, w[k].maxRef); t.unequal += w[k].unequal; t.nan += w[k].nan; }
        partials[blockIdx.x] = t;
    }
}

static int compareImpl(dfsa_state* a, const double2* b, bool hash, uint64_t seed, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef) {
    DFSA_REQUIRE(maxAbsDiff && numUnequal, "null output");
    const unsigned grid = dfsaGrid(a->numAmps, 256, 4, 8);
    double2* scratch;
    DFSA_TRY(dfsaScratch(grid * sizeof(CmpPartial), &scratch));
    CmpPartial* dev = (CmpPartial*)scratch;
    const uint64_t first = (uint64_t)a->rank << a->logNumAmps;
    if (hash) compareKernel<true><<<grid, 256, 0, g_ctx.compute>>>(a->arr[0], nullptr, a->numAmps, seed, first, dev);
    else compareKernel<false><<<grid, 256, 0, g_ctx.compute>>>(a->arr[0], b, a->numAmps, 0, first, dev);
    DFSA_LAUNCH_CHECK();
    std::vector<CmpPartial> host(grid);
    DFSA_CUDA(cudaMemcpyAsync(host.data(), dev, grid * sizeof(CmpPartial), cudaMemcpyDeviceToHost, g_ctx.compute));
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.compute));
    double mx[2] = {0.0, 0.0}, sums[2] = {0.0, 0.0};
    for (const CmpPartial& p : host) { mx[0] = fmax(mx[0], p.maxDiff); mx[1] = fmax(mx[1], p.maxRef); sums[0] += p.unequal; sums[1] += p.nan; }
    if (g_ctx.size > 1) {
        DFSA_TRY(dfsaAllreduceDoubles(mx, 2, true));
        DFSA_TRY(dfsaAllreduceDoubles(sums, 2, false));
    }
    *maxAbsDiff = sums[1] > 0.0 ? nan("") : mx[0];
    *numUnequal = (uint64_t)sums[0];
    if (maxAbsRef) *maxAbsRef = mx[1];
    return DFSA_OK;
}

extern "C" int dfsa_state_compare(dfsa_state* a, dfsa_state* b, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(a && b && a->numAmps == b->numAmps && a->isDensity == b->isDensity, "states must have the same shape");
    return compareImpl(a, b->arr[0], false, 0, maxAbsDiff, numUnequal, maxAbsRef);
}

extern "C" int dfsa_state_compare_hash(dfsa_state* s, uint64_t seed, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    return compareImpl(s, nullptr, true, seed, maxAbsDiff, numUnequal, maxAbsRef);
}

// sum |amp|^2 : block partials -> host sum (deterministic order)
__global__ void __launch_bounds__(256) norm2Kernel(const double2* __restrict__ amps, uint64_t n, double* partials) {
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double2 a = amps[i];
        acc = fma(a.x, a.x, fma(a.y, a.y, acc));
    }
    __shared__ double warpSums[8];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += warpSums[w];
        partials[blockIdx.x] = t;
    }
}

extern "C" int dfsa_state_norm2(dfsa_state* s, double* out) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && out, "null argument");
    unsigned grid = dfsaGrid(s->numAmps, 256, 4, 8);
    double2* scratch;
    DFSA_TRY(dfsaScratch(grid * sizeof(double), &scratch));
    norm2Kernel<<<grid, 256, 0, g_ctx.compute>>>(s->arr[0], s->numAmps, (double*)scratch);
    DFSA_LAUNCH_CHECK();
    std::vector<double> partials(grid);
    DFSA_CUDA(cudaMemcpyAsync(partials.data(), scratch, grid * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.compute));
    DFSA_CUDA(cudaStreamSynchronize(g_ctx.compute));
    double reim[2] = {0.0, 0.0};
    for (double p : partials) reim[0] += p;
    if (g_ctx.size > 1) DFSA_TRY(dfsa_x_allreduce_amp(reim));
    *out = reim[0];
    return DFSA_OK;
}
