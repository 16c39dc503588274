// dfsa_stream_kernels.cuh -- the one streaming-kernel skeleton every HBM-bound loop of the path uses.
//
// All of K1-K3, K5-K16, K18-K22 are "touch each amplitude (pair / quad) once": no reuse, so no
// shared memory; the job is to keep enough 128-bit loads in flight to cover HBM latency. Each thread issues
// UNROLL independent item loads (an item = 1, 2 or 4 amplitudes, all 16-byte ld.global.v2.f64) before any
// arithmetic or store; consecutive threads take consecutive items so every warp-level access is a run of
// contiguous 16-byte amplitudes; blocks walk the item space grid-stride with a grid sized as a multiple of
// the SM count.  Loads and stores are passed as two device lambdas so that the in-place read-modify-write
// cannot serialise the loads of item u+1 behind the stores of item u.
#pragma once
#include <stdlib.h>
#include "dfsa_internal.cuh"

constexpr int DFSA_TPB = 256;

template <int UNROLL, class T, class LD, class ST>
__global__ void __launch_bounds__(DFSA_TPB) streamKernel(uint64_t numItems, LD ld, ST st) {
    const uint64_t chunk = (uint64_t)DFSA_TPB * UNROLL;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < numItems; base += (uint64_t)gridDim.x * chunk) {
        T v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            uint64_t j = base + (uint64_t)u * DFSA_TPB + threadIdx.x;
            if (j < numItems) v[u] = ld(j);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            uint64_t j = base + (uint64_t)u * DFSA_TPB + threadIdx.x;
            if (j < numItems) st(j, v[u]);
        }
    }
}

// How much to keep in flight: measured on B200 (tools/onetarg_sweep.py, tools/microbench_cu/rmw_variants.cu,
// profiles/r01_rmw_variants.jsonl) the in-place read-modify-write stream is fastest with about 2048 outstanding 16-byte
// loads per SM spread over MANY threads (1024 threads x 1 pair: 6.18-6.34 TB/s for every target position) and gets
// slower both with fewer (latency-bound) and with more (full occupancy: 5.8-6.0 TB/s -- more concurrent streams, fewer
// open-page hits). So the persistent grid is SMs x B with B = inflight / (256 x loads per thread), capped by occupancy;
// DFSA_STREAM_INFLIGHT (default 2048) overrides the target.
//
// REMOTE = some of the loads go to a peer GPU over NVLink (the fused exchange kernels): latency is a few microseconds instead
// of ~1, so the same bandwidth-delay argument asks for several times as many loads in flight; measured on 2 B200s
// (tools/link_sweep.py, profiles/r02_link_sweep_n2.jsonl). DFSA_REMOTE_INFLIGHT (default 8192 per SM, local + remote) overrides.
template <int UNROLL, class T, class LD, class ST, bool REMOTE = false>
static int launchStream(uint64_t numItems, LD ld, ST st) {
    if (numItems == 0) return DFSA_OK;
    static int blocksPerSM = 0;
    if (blocksPerSM == 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, streamKernel<UNROLL, T, LD, ST>, DFSA_TPB, 0) != cudaSuccess || occ < 1) occ = 4;
        const char* e = getenv(REMOTE ? "DFSA_REMOTE_INFLIGHT" : "DFSA_STREAM_INFLIGHT");
        const int inflight = e ? atoi(e) : (REMOTE ? 8192 : 2048);
        const int loadsPerThread = UNROLL * T::kLoads;
        int want = inflight / (DFSA_TPB * loadsPerThread);
        if (want < 1) want = 1;
        blocksPerSM = want < occ ? want : occ;
    }
    unsigned grid = dfsaGrid(numItems, DFSA_TPB, UNROLL, (unsigned)blocksPerSM);
    streamKernel<UNROLL, T, LD, ST><<<grid, DFSA_TPB, 0, dfsaCtx().compute>>>(numItems, ld, st);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

// fused exchange kernels: two items per thread, sized for NVLink latency
template <class T, class LD, class ST>
static int launchStreamRemote(uint64_t numItems, LD ld, ST st) { return launchStream<2, T, LD, ST, true>(numItems, ld, st); }

// item types; kLoads = 16-byte loads one item keeps in flight (used to size the grid, see launchStream)
struct Amp1 { static constexpr int kLoads = 1; double2 a; };
struct Amp2 { static constexpr int kLoads = 2; double2 a0, a1; };
struct Amp4 { static constexpr int kLoads = 4; double2 a00, a01, a10, a11; };
struct Amp3 { static constexpr int kLoads = 3; double2 a, b, c; };
struct PairAt { static constexpr int kLoads = 2; double2 a0, a1; uint64_t idx; };   // an amplitude pair plus the index it was loaded from
struct QuadAt { static constexpr int kLoads = 4; double2 v0, v1, v2, v3; uint64_t idx; };   // four amplitudes spanned by two target bits
// a pair that is only MOVED (swap): half the traffic per item of a read-modify-write pair, so twice as many in flight
struct MovePair { static constexpr int kLoads = 1; double2 a0, a1; uint64_t idx; };
