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

// blocksPerSM: 2048 resident threads / 256 = 8 when registers allow; the kernels here stay under 32 regs*... the
// launch helper asks the runtime so the grid is always SMs x resident blocks.
template <int UNROLL, class T, class LD, class ST>
static int launchStream(uint64_t numItems, LD ld, ST st) {
    if (numItems == 0) return DFSA_OK;
    static int blocksPerSM = 0;
    if (blocksPerSM == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, streamKernel<UNROLL, T, LD, ST>, DFSA_TPB, 0) != cudaSuccess || blocksPerSM < 1)
            blocksPerSM = 4;
    }
    unsigned grid = dfsaGrid(numItems, DFSA_TPB, UNROLL, (unsigned)blocksPerSM);
    streamKernel<UNROLL, T, LD, ST><<<grid, DFSA_TPB, 0, dfsaCtx().compute>>>(numItems, ld, st);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

struct Amp1 { double2 a; };
struct Amp2 { double2 a0, a1; };
struct Amp4 { double2 a00, a01, a10, a11; };
struct Amp3 { double2 a, b, c; };
struct PairAt { double2 a0, a1; uint64_t idx; };   // an amplitude pair plus the index it was loaded from
