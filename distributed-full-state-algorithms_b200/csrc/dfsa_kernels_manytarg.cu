// dfsa_kernels_manytarg.cu -- K4: the dense 2^t x 2^t gate of local_statevector.hpp:72-99.
// Row/column bit i of the gate belongs to targets[i] in the CALLER's order; amplitudes outside the targets are
// enumerated by inserting zeros at the sorted target positions. 32*A bytes of HBM traffic and 8*2^t flop per
// amplitude (6*2^t in the 3M form used here): HBM-bound for t <= 4, FP64-pipe-bound from t = 5.
//
// Kernels:
//   t == 1   the pair-stream kernel of K1 (same operator)
//   t == 2   register-resident 4x4 matvec on the streaming skeleton (quad items, gate in kernel parameters); HBM-bound
//   t = 3..6 manyTargSpecKernel<T>: FP64 tensor cores (mma.sync m16n8k16.f64 -> SASS DMMA.8x8x4), warp-specialised:
//            mover warps stream 512-amplitude tiles HBM -> shared memory -> HBM (cp.async + mbarriers), compute warps do
//            nothing but LDS -> DMMA -> STS with the gate held as A-fragments in registers.
//            History (profiles/r01*_ncu_manytarg*): the DFMA version of t=5 was issue-bound at 18 % FP64-pipe utilisation
//            and 8 % of DRAM bandwidth, i.e. compute- not HBM-bound -- the case north_star reserves the tensor path for;
//            one warp per tile with per-element address tables reached 54 %, then 81 % of the pipe in 4M form; the
//            3M form with the rows split over a warp pair 70 %; this kernel 81 % in 3M form (25 % fewer flops).
//            t == 6 (krausMap / density-matrix gates on 3 qubits): same kernel, the 64x64 gate's A-fragments in shared memory.
//   t = 7..11 manyTargGemmKernel: tiled FP64 DMMA GEMM over the groups, gate streamed from L2 as A-fragments (krausMap on 4-5 qubits)
//   t >= 12  manyTargGenericKernel   one block per 2^t group, gate streamed from L2, warp-per-row reduction
//            (also: t = 3..6 on shards smaller than one tile)
#include <algorithm>
#include <vector>

#include <string.h>
#include "dfsa_stream_kernels.cuh"

// ---------------------------------------------------------------------------------------------------------
// FP64 tensor-core MMA. Fragment maps (m16n8k16.row.col.f64; g = lane/4, q = lane%4):
//   a[v]: row g + 8(v&1),  col q + 4(v>>1)      b[v]: k = q + 4v, n = g      c[v]: row g + 8(v>>1), col 2q + (v&1)
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mbarriers in shared memory (full / done / drained hand-over between mover and compute warps)
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
// arrive-on triggered when all cp.async issued so far by this thread have landed (counted in the barrier's init count)
__device__ __forceinline__ void cpAsyncArriveNoinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void ldsAmp(double& re, double& im, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(re), "=d"(im) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsAmp(unsigned addr, double re, double im) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(re), "d"(im) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Geometry. A tile = the 512 amplitudes (8 KiB) spanned by the t target bits and the F = 9 - t lowest non-target bits;
// it is staged as a slab X[row][col]: row = gate-ordered target bits, col = free bits (2^F vectors the same gate acts on).
// A block runs 4 tile STREAMS; each stream has one mover warp and two compute warps, NIN input slabs and NOUT output
// slabs (4 + 2; t = 6: 2 + 1), handed over through mbarriers:
//   full[s]    mover -> compute: tile landed in input slab s (cp.async.mbarrier.arrive.noinc of the 32 mover lanes)
//   done[o]    compute -> mover: output slab o holds a tile's results and its input slab is free again (both warps arrive)
//   drained[o] mover -> compute: output slab o has been read out, it may be overwritten
// Tile i of a stream uses input slab i % NIN and output slab i % NOUT; the mover loads tile i + NIN as soon as tile i is done,
// i.e. NIN - 1 tile-times ahead of its use. Why specialise: with load/store and DMMA phases in the same warps, the warps of
// a scheduler share the tensor pipe fairly and drift into lock-step -- all computing, then all moving data with the pipe idle.
//
// The two compute warps of a stream split the tile: t = 5, 6 by gate ROWS (t = 5: 16 each, the three A-matrices of the 3M
// product then cost 96 registers; t = 6: 32 each, A-fragments in shared memory), t = 3, 4 by COLUMNS (the whole gate fits:
// 16 / 48 registers).
//
// Slab layout: a slab is the tile in ADDRESS order -- tile element e (bit p of e = p-th lowest tile bit of the shard index)
// sits at byte offset XOR_p bit_p(e) * bitOff[p], bitOff[p] = (16 << p) ^ (m[p] << 4) with a 3-bit swizzle m[p] for p >= 3.
// So the mover copies contiguous runs to contiguous runs: the two 16-byte halves of a 32-byte sector stay neighbours (ncu:
// with a row-major [gate row][vector] slab, a target on index bit 0 sent them to different rows and cp.async fetched every
// sector from L2 twice -- 2.7x the L2 read sectors, 4.0 instead of 5.5 TB/s), and its accesses are bank-conflict free by
// construction. The transposition into gate rows / vectors happens in the compute warps' LDS / STS addresses instead, where
// it is free: gate-row bit i and vector bit j are tile bits, so offset(row, col) = XOR of per-bit constants (SpecLayout),
// every address is slab ^ (thread part) ^ (instruction part), one LOP3. The swizzle m[] is chosen per launch (host, brute
// force over <= 7^5 candidates) so that the eight lanes of a quarter-warp hit eight different 16-byte bank groups in both
// fragment patterns: B reads vary (row bit 0, row bit 1, col bit 0), result writes vary (col bit 1, col bit 2, row bit 0).
template <int T> struct SpecGeom {
    static constexpr unsigned D = 1u << T, TILE_BITS = 9, F = TILE_BITS - T, COLS = 1u << F;
    // t = 6: the 64x64 gate cannot live in registers; its A-fragments (three 3M matrices, 96 KiB) sit in shared memory,
    // which leaves room for 2 + 1 slabs per stream -- enough, the tile is FP64-bound four times over
    static constexpr unsigned STREAMS = 4, NIN = (T == 6) ? 2 : 4, NOUT = (T == 6) ? 1 : 2;
    static constexpr unsigned SLAB_BYTES = 16u << TILE_BITS;                   // 8 KiB
    static constexpr unsigned THREADS = 32 * (2 * STREAMS + STREAMS);          // 8 compute warps + 4 mover warps
    static constexpr unsigned GATE_BYTES = 3u * (D / 16) * (D / 16) * 2048u;   // [matrix][mb][kb] A-fragments of 2 KiB, when they live in shared memory
    static constexpr size_t smemBytes(bool smemA) { return (size_t)STREAMS * (NIN + NOUT) * SLAB_BYTES + (smemA ? GATE_BYTES : 0u) + SLAB_BYTES; }   // + alignment slack
};
struct SpecLayout { uint32_t rowBit[6], colBit[6]; };               // slab byte-offset contribution of gate-row bit i / vector bit j
// i part of tile element (lane | i << 5): byte offset in the shard and in the slab. Rides in the kernel parameters, so after
// unrolling every entry is a constant-bank operand (the first tensor kernel read two shared-memory tables and spent ~15 ALU
// instructions per 16-byte element: 1200 non-DMMA instructions per tile against 512 DMMA slots).
struct SpecMap { uint64_t gByte[16]; uint32_t sByte[16]; };

template <int T>
__host__ __device__ __forceinline__ unsigned specOffset(unsigned row, unsigned col, const SpecLayout& z) {
    unsigned o = 0;
#pragma unroll
    for (int i = 0; i < T; i++) o ^= ((row >> i) & 1u) ? z.rowBit[i] : 0u;
#pragma unroll
    for (int j = 0; j < (int)SpecGeom<T>::F; j++) o ^= ((col >> j) & 1u) ? z.colBit[j] : 0u;
    return o;
}

template <int T, bool SMEM_A>
__global__ void __launch_bounds__(SpecGeom<T>::THREADS, 1)
manyTargSpecKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec, BitSpec localPos, const double2* __restrict__ gate, SpecMap map, SpecLayout lay) {
    using G = SpecGeom<T>;
    constexpr unsigned D = G::D, STREAMS = G::STREAMS, NIN = G::NIN, NOUT = G::NOUT, SLAB_BYTES = G::SLAB_BYTES, TILE_BITS = G::TILE_BITS;
    extern __shared__ double2 smem[];
    __shared__ uint64_t full[STREAMS][NIN], done[STREAMS][NOUT], drained[STREAMS][NOUT];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool mover = warp >= 2 * STREAMS;
    const unsigned stream = mover ? warp - 2 * STREAMS : warp >> 1, h = warp & 1u;
    // shared-window addresses of this stream's slabs (8 KiB-aligned)
    const unsigned slab0 = ((smemAddr(smem) + SLAB_BYTES - 1u) & ~(SLAB_BYTES - 1u)) + stream * ((NIN + NOUT) * SLAB_BYTES);
    const unsigned in0 = slab0, out0 = slab0 + NIN * SLAB_BYTES;

    if (threadIdx.x < STREAMS) {
        for (unsigned s = 0; s < NIN; s++) mbarInit(&full[threadIdx.x][s], 32);       // the mover's 32 lanes
        for (unsigned o = 0; o < NOUT; o++) { mbarInit(&done[threadIdx.x][o], 2); mbarInit(&drained[threadIdx.x][o], 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // SMEM_A: A-fragments of the three 3M matrices, [matrix m][mb][kb][chunk c][lane] as 16-byte (a[2c], a[2c+1]) pairs, so
    // that a warp's LDS.128 of one chunk is 512 contiguous bytes
    constexpr unsigned MBT = D / 16;                                  // 16-row / 16-column blocks of the gate
    const unsigned gateS = ((smemAddr(smem) + SLAB_BYTES - 1u) & ~(SLAB_BYTES - 1u)) + STREAMS * ((NIN + NOUT) * SLAB_BYTES);
    if constexpr (SMEM_A) {
        for (unsigned idx = threadIdx.x; idx < 3u * MBT * MBT * 4u * 32u; idx += G::THREADS) {
            const unsigned ln = idx & 31u, c = (idx >> 5) & 3u, blk = idx >> 7, kb = blk % MBT, mb = (blk / MBT) % MBT, m = blk / (MBT * MBT);
            double val[2];
#pragma unroll
            for (unsigned w = 0; w < 2; w++) {
                const unsigned v = 2 * c + w;                        // a[v]: row g + 8(v&1), col q + 4(v>>1)
                const double2 e = gate[(16 * mb + (ln >> 2) + 8 * (v & 1)) * D + 16 * kb + (ln & 3u) + 4 * (v >> 1)];
                val[w] = m == 0 ? e.x : (m == 1 ? e.y - e.x : e.x + e.y);
            }
            stsAmp(gateS + idx * 16u, val[0], val[1]);
        }
    }
    __syncthreads();

    const uint64_t stride = (uint64_t)gridDim.x * STREAMS;
    const uint64_t tile0 = (uint64_t)blockIdx.x * STREAMS + stream;
    const unsigned numMine = tile0 < numTiles ? (unsigned)((numTiles - tile0 + stride - 1) / stride) : 0u;

    if (mover) {
        // element (lane | i << 5), i < 16: lane part in registers, i part = map; slab offsets combine by XOR
        uint64_t laneOff = 0;
        unsigned laneRow = 0, laneCol = 0;
#pragma unroll
        for (unsigned b = 0; b < 5; b++) {
            const unsigned bit = (lane >> b) & 1u, role = localPos.pos[b];
            laneOff |= (uint64_t)bit << tileSpec.pos[b];
            if (role < (unsigned)T) laneRow |= bit << role; else laneCol |= bit << (role - T);
        }
        char* laneG = reinterpret_cast<char*>(amps) + (laneOff << 4);
        const unsigned laneS = specOffset<T>(laneRow, laneCol, lay);
        auto load = [&](unsigned i) {                                  // tile i of this stream -> input slab i % NIN
            const char* src = laneG + (insertZeroBitsN<TILE_BITS>(tile0 + (uint64_t)i * stride, tileSpec) << 4);
            const unsigned s = i % NIN, dst = (in0 + s * SLAB_BYTES) ^ laneS;
#pragma unroll
            for (unsigned e = 0; e < 16; e++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst ^ map.sByte[e]), "l"(src + map.gByte[e]) : "memory");
            cpAsyncArriveNoinc(&full[stream][s]);
        };
        for (unsigned i = 0; i < NIN && i < numMine; i++) load(i);
        for (unsigned i = 0; i < numMine; i++) {
            const unsigned o = i % NOUT;
            mbarWait(&done[stream][o], (i / NOUT) & 1u);
            char* dst = laneG + (insertZeroBitsN<TILE_BITS>(tile0 + (uint64_t)i * stride, tileSpec) << 4);
            const unsigned src = (out0 + o * SLAB_BYTES) ^ laneS;
#pragma unroll
            for (unsigned e0 = 0; e0 < 16; e0 += 8) {
                double2 v[8];
#pragma unroll
                for (unsigned e = 0; e < 8; e++) ldsAmp(v[e].x, v[e].y, src ^ map.sByte[e0 + e]);
#pragma unroll
                for (unsigned e = 0; e < 8; e++) *reinterpret_cast<double2*>(dst + map.gByte[e0 + e]) = v[e];
            }
            __syncwarp();                                               // every lane holds its share of the slab in registers
            if (lane == 0) mbarArrive(&drained[stream][o]);
            if (i + NIN < numMine) load(i + NIN);                       // input slab i % NIN was released by done
        }
        return;
    }

    // ---- compute warps
    const unsigned g = lane >> 2, q = lane & 3u;
    // thread parts of the operand (row q, column g) and result (row g [+ 16h], column 2q) slab offsets
    const unsigned laneB = specOffset<T>(q, g, lay);
    const unsigned laneC = specOffset<T>((T >= 5 ? (D / 2) * h : 0) + g, 2 * q, lay);
    // the gate as A-fragments. 3M complex product (t = 4, 5): with k1 = G_re (X_re + X_im), k2 = (G_im - G_re) X_re,
    // k3 = (G_re + G_im) X_im:  Y_re = k1 - k3, Y_im = k1 + k2 -- three real MMAs where the 4M form needs four.
    // t = 3: the 8x8 complex gate as ONE real 16x16 matrix [[G_re, -G_im], [G_im, G_re]] acting on [X_re; X_im].
    constexpr int KB = (T == 5) ? 2 : 1;                            // 16-column blocks of the gate (register-resident forms)
    double ar[KB][8], ad[KB][8], as[KB][8];
#pragma unroll
    for (int kb = 0; kb < (SMEM_A ? 0 : KB); kb++)
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const unsigned row = g + 8 * (v & 1), col = q + 4 * (v >> 1);
            if constexpr (T == 3) {
                const double2 e = gate[(row & 7u) * D + (col & 7u)];
                ar[kb][v] = ((row < 8) == (col < 8)) ? e.x : ((row < 8) ? -e.y : e.y);
                ad[kb][v] = 0.0; as[kb][v] = 0.0;
            } else {
                const double2 e = gate[((T == 5 ? 16 * h : 0) + row) * D + 16 * kb + col];
                ar[kb][v] = e.x;
                ad[kb][v] = e.y - e.x;
                as[kb][v] = e.x + e.y;
            }
        }
    // column blocks (8 vectors each) this warp works on: all of them for t = 5 (rows are split), half of them otherwise
    constexpr unsigned NBW = SMEM_A ? 1 : ((T >= 5) ? G::COLS / 8 : G::COLS / 16);
    const unsigned nb0 = (T >= 5) ? 0u : h * NBW;

    for (unsigned i = 0; i < numMine; i++) {
        const unsigned s = i % NIN, o = i % NOUT;
        const unsigned xB = (in0 + s * SLAB_BYTES) ^ laneB, yC = (out0 + o * SLAB_BYTES) ^ laneC;
        mbarWait(&full[stream][s], (i / NIN) & 1u);
        if (i >= NOUT) mbarWait(&drained[stream][o], ((i / NOUT) - 1u) & 1u);
#pragma unroll
        for (unsigned nbi = 0; nbi < NBW; nbi++) {
            const unsigned nb = nb0 + nbi;
            if constexpr (SMEM_A) {
                // this warp: gate rows (D/2)h .. (D/2)h + D/2 - 1 (MBW 16-row blocks), all MBT 16-column blocks, every column
                // block of the tile at once (NBJ: one A-fragment load feeds NBJ MMAs); A-fragments from shared memory
                constexpr unsigned MBW = MBT / 2, NBJ = G::COLS / 8;
                double k1[NBJ][MBW][4], k2[NBJ][MBW][4], k3[NBJ][MBW][4];
#pragma unroll
                for (unsigned j = 0; j < NBJ; j++)
#pragma unroll
                    for (unsigned mbl = 0; mbl < MBW; mbl++)
#pragma unroll
                        for (int v = 0; v < 4; v++) { k1[j][mbl][v] = 0.0; k2[j][mbl][v] = 0.0; k3[j][mbl][v] = 0.0; }
                const unsigned laneA = gateS + (h * (MBW * MBT * 4u * 32u) + lane) * 16u;   // + ((((m*MBT + mbl)*MBT + kb)*4 + c) << 9)
#pragma unroll
                for (unsigned kb = 0; kb < MBT; kb++) {
                    double xr[NBJ][4], xi[NBJ][4], xs[NBJ][4];
#pragma unroll
                    for (unsigned j = 0; j < NBJ; j++)
#pragma unroll
                        for (unsigned v = 0; v < 4; v++) {
                            ldsAmp(xr[j][v], xi[j][v], xB ^ specOffset<T>(16 * kb + 4 * v, j * 8, lay));
                            xs[j][v] = xr[j][v] + xi[j][v];
                        }
#pragma unroll
                    for (unsigned mbl = 0; mbl < MBW; mbl++) {
                        // matrix order 1, 2, 0: the product that needs the DADD results (x_re + x_im) goes last
#pragma unroll
                        for (unsigned mi = 0; mi < 3; mi++) {
                            const unsigned m = (mi + 1) % 3;
                            double a[8];
#pragma unroll
                            for (unsigned c = 0; c < 4; c++)
                                ldsAmp(a[2 * c], a[2 * c + 1], laneA + ((((m * MBT + mbl) * MBT + kb) * 4 + c) << 9));
#pragma unroll
                            for (unsigned j = 0; j < NBJ; j++) {
                                if (m == 1) dmma16816(k2[j][mbl], a, xr[j]);
                                else if (m == 2) dmma16816(k3[j][mbl], a, xi[j]);
                                else dmma16816(k1[j][mbl], a, xs[j]);
                            }
                        }
                    }
                }
#pragma unroll
                for (unsigned j = 0; j < NBJ; j++)
#pragma unroll
                    for (unsigned mbl = 0; mbl < MBW; mbl++)
#pragma unroll
                        for (unsigned v = 0; v < 4; v++)                // c[v]: row (D/2)h + 16mbl + g + 8(v>>1), column 8j + 2q + (v&1)
                            stsAmp(yC ^ specOffset<T>(16 * mbl + 8 * (v >> 1), j * 8 + (v & 1), lay), k1[j][mbl][v] - k3[j][mbl][v], k1[j][mbl][v] + k2[j][mbl][v]);
                (void)nb;
            } else if constexpr (T == 3) {
                // b[v]: k = q + 4v; k < 8 -> X_re row k, k >= 8 -> X_im row k - 8
                double x0r, x0i, x1r, x1i;
                ldsAmp(x0r, x0i, xB ^ specOffset<T>(0, nb * 8, lay));
                ldsAmp(x1r, x1i, xB ^ specOffset<T>(4, nb * 8, lay));
                const double bfrag[4] = {x0r, x1r, x0i, x1i};
                double c[4] = {0.0, 0.0, 0.0, 0.0};
                dmma16816(c, ar[0], bfrag);
                // c[v]: row g + 8(v>>1) (rows 8..15 = imaginary parts of complex row g), col 2q + (v&1)
                stsAmp(yC ^ specOffset<T>(0, nb * 8, lay), c[0], c[2]);
                stsAmp(yC ^ specOffset<T>(0, nb * 8 + 1, lay), c[1], c[3]);
            } else {
                double k1[4] = {0.0, 0.0, 0.0, 0.0}, k2[4] = {0.0, 0.0, 0.0, 0.0}, k3[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    double xr[4], xi[4], xs[4];
#pragma unroll
                    for (unsigned v = 0; v < 4; v++) {                  // b[v]: k = q + 4v (row 16kb + q + 4v), n = nb*8 + g
                        ldsAmp(xr[v], xi[v], xB ^ specOffset<T>(16 * kb + 4 * v, nb * 8, lay));
                        xs[v] = xr[v] + xi[v];
                    }
                    // the product that needs the DADD results goes last: the adds retire behind the other two products' DMMAs
                    dmma16816(k2, ad[kb], xr);
                    dmma16816(k3, as[kb], xi);
                    dmma16816(k1, ar[kb], xs);
                }
#pragma unroll
                for (unsigned v = 0; v < 4; v++)                        // c[v]: row g + 8(v>>1) [+ 16h], column nb*8 + 2q + (v&1)
                    stsAmp(yC ^ specOffset<T>(8 * (v >> 1), nb * 8 + (v & 1), lay), k1[v] - k3[v], k1[v] + k2[v]);
            }
        }
        __syncwarp();                                                   // all of this warp's reads and writes of the slabs are done
        if (lane == 0) mbarArrive(&done[stream][o]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// t >= 7 (krausMap superoperators) and small shards: one block per 2^t-amplitude group staged in shared memory,
// gate rows streamed from global memory (L2-resident), a warp per output row, shuffle reduction over columns.
__global__ void __launch_bounds__(256) manyTargGenericKernel(double2* amps, uint64_t numGroups, BitSpec sortedTargs, BitSpec callerTargs,
                                                            unsigned t, const double2* __restrict__ gate, double2* __restrict__ outStage) {
    extern __shared__ double2 smem[];
    const uint64_t d = 1ULL << t;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, numWarps = blockDim.x >> 5;
    double2* out = outStage + (size_t)blockIdx.x * d;           // per-block result slab (results must not overwrite inputs early)
    for (uint64_t grp = blockIdx.x; grp < numGroups; grp += gridDim.x) {
        const uint64_t base = insertZeroBits(grp, sortedTargs);
        __syncthreads();
        for (uint64_t l = threadIdx.x; l < d; l += blockDim.x) {
            uint64_t g = base;
            for (unsigned b = 0; b < t; b++) g |= ((l >> b) & 1ULL) << callerTargs.pos[b];
            smem[l] = amps[g];
        }
        __syncthreads();
        for (uint64_t r = warp; r < d; r += numWarps) {
            double2 acc = make_double2(0.0, 0.0);
            const double2* grow = gate + r * d;
            for (uint64_t l = lane; l < d; l += 32) acc = cfma(grow[l], smem[l], acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            }
            if (lane == 0) out[r] = acc;
        }
        __syncthreads();
        for (uint64_t l = threadIdx.x; l < d; l += blockDim.x) {
            uint64_t g = base;
            for (unsigned b = 0; b < t; b++) g |= ((l >> b) & 1ULL) << callerTargs.pos[b];
            amps[g] = out[l];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// t = 7 .. 11 (krausMap superoperators on 4-5 qubits): the gate no longer fits on chip, so the operation is
// run as what it is -- a GEMM  Y[D x M] = G[D x D] X[D x M]  over the M = A / D groups of amplitudes -- on the FP64 tensor cores,
// in the 3M form of the complex product, with the gate STREAMED: a prep kernel turns it into the three real matrices as
// mma A-fragments (fragment-major, so a warp's load of one fragment is 4 x 512 contiguous bytes), which then come from L2
// (<= 96 MiB, resident) straight into registers. Block tile = 128 gate rows (8 warps x 16) x 64 groups; K loop over 16-row
// slices of X, double-buffered in shared memory with cp.async (XOR-swizzled so the B-fragment reads are conflict free);
// accumulators (3 x 8 x 4 doubles per thread) stay in registers for the whole K loop. One A-fragment load feeds 8 MMAs:
// 16 flop per byte from L2, ~2.3 TB/s of L2 traffic at the FP64 peak. Results go to a scratch slab [row][group] and are
// copied back by a second pass (other blocks still need the old amplitudes of the same groups), chunk by chunk so the
// scratch stays at 256 MiB whatever the shard size.
constexpr unsigned GEMM_BM = 128, GEMM_BN = 64, GEMM_THREADS = 256;

__global__ void __launch_bounds__(256) gemmPrepKernel(const double2* __restrict__ gate, unsigned D, double2* __restrict__ afrag) {
    const unsigned MBT = D / 16;
    const uint64_t total = 3ull * MBT * MBT * 4u * 32u;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned ln = idx & 31u, c = (idx >> 5) & 3u;
        const uint64_t blk = idx >> 7;
        const unsigned kb = blk % MBT, mb = (blk / MBT) % MBT, m = (unsigned)(blk / ((uint64_t)MBT * MBT));
        double val[2];
#pragma unroll
        for (unsigned w = 0; w < 2; w++) {
            const unsigned v = 2 * c + w;                            // a[v]: row g + 8(v&1), col q + 4(v>>1)
            const double2 e = gate[(uint64_t)(16 * mb + (ln >> 2) + 8 * (v & 1)) * D + 16 * kb + (ln & 3u) + 4 * (v >> 1)];
            val[w] = m == 0 ? e.x : (m == 1 ? e.y - e.x : e.x + e.y);
        }
        afrag[idx] = make_double2(val[0], val[1]);
    }
}

// ROWS_FAST: gate bit 0 sits on index bit 0, so the 16 rows of a slice are runs of contiguous amplitudes while neighbouring
// groups are 2^t-strided: the loader's lanes then walk rows first (otherwise groups first, which are contiguous when the
// low index bits are not targets).
template <bool ROWS_FAST>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
manyTargGemmKernel(const double2* __restrict__ amps, double2* __restrict__ ytmp, uint64_t group0, unsigned chunkGroups, BitSpec sortedTargs,
                   const uint64_t* __restrict__ rowOff, unsigned D, const double2* __restrict__ afrag) {
    __shared__ __align__(16) double2 xs[2][16 * GEMM_BN];             // [buffer][k][n ^ ((k & 3) << 1)]
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3u;
    const unsigned MBT = D / 16, col0 = blockIdx.x * GEMM_BN, mb = blockIdx.y * (GEMM_BM / 16) + warp;
    // loader role: 4 of the 16 x 64 amplitudes of every slice -- (row kq + 4j, column n), or with ROWS_FAST (row k, column nq + 16j)
    const unsigned n = ROWS_FAST ? (threadIdx.x >> 4) : (threadIdx.x & (GEMM_BN - 1)), kq = ROWS_FAST ? (threadIdx.x & 15u) : (threadIdx.x / GEMM_BN);
    const double2* colBase[ROWS_FAST ? 4 : 1];
    bool colValid[ROWS_FAST ? 4 : 1];
#pragma unroll
    for (unsigned j = 0; j < (ROWS_FAST ? 4u : 1u); j++) {
        const unsigned col = col0 + n + 16 * j;
        colValid[j] = col < chunkGroups;
        colBase[j] = amps + (colValid[j] ? insertZeroBits(group0 + col, sortedTargs) : 0ull);
    }
    auto loadSlice = [&](unsigned kb, unsigned buf) {
#pragma unroll
        for (unsigned j = 0; j < 4; j++) {
            const unsigned k = ROWS_FAST ? kq : kq + 4 * j, nn = ROWS_FAST ? n + 16 * j : n, cj = ROWS_FAST ? j : 0;
            double2* dst = &xs[buf][k * GEMM_BN + (nn ^ ((k & 3u) << 1))];
            if (colValid[cj]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr(dst)), "l"(colBase[cj] + __ldg(&rowOff[16 * kb + k])) : "memory");
            else *dst = make_double2(0.0, 0.0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto loadA = [&](unsigned kb, double (&ar)[8], double (&ad)[8], double (&as)[8]) {
        const double2* base = afrag + ((uint64_t)mb * MBT + kb) * 128u + lane;
        const uint64_t mStride = (uint64_t)MBT * MBT * 128u;
#pragma unroll
        for (unsigned c = 0; c < 4; c++) {
            const double2 r0 = __ldg(base + c * 32u), r1 = __ldg(base + mStride + c * 32u), r2 = __ldg(base + 2 * mStride + c * 32u);
            ar[2 * c] = r0.x; ar[2 * c + 1] = r0.y;
            ad[2 * c] = r1.x; ad[2 * c + 1] = r1.y;
            as[2 * c] = r2.x; as[2 * c + 1] = r2.y;
        }
    };
    double k1[8][4], k2[8][4], k3[8][4];
#pragma unroll
    for (unsigned nb = 0; nb < 8; nb++)
#pragma unroll
        for (unsigned v = 0; v < 4; v++) { k1[nb][v] = 0.0; k2[nb][v] = 0.0; k3[nb][v] = 0.0; }

    loadSlice(0, 0);
    for (unsigned kb = 0; kb < MBT; kb++) {
        const unsigned buf = kb & 1u;
        double ar[8], ad[8], as[8];
        loadA(kb, ar, ad, as);                                        // L2 -> registers, in flight while the slice lands
        if (kb + 1 < MBT) { loadSlice(kb + 1, buf ^ 1u); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const unsigned xB = smemAddr(&xs[buf][0]);
#pragma unroll
        for (unsigned nb = 0; nb < 8; nb++) {
            double xr[4], xi[4], xsum[4];
#pragma unroll
            for (unsigned v = 0; v < 4; v++) {                        // b[v]: k = q + 4v, n = nb * 8 + g
                ldsAmp(xr[v], xi[v], xB + (((q + 4 * v) * GEMM_BN + ((nb * 8 + g) ^ (q << 1))) << 4));
                xsum[v] = xr[v] + xi[v];
            }
            dmma16816(k2[nb], ad, xr);
            dmma16816(k3[nb], as, xi);
            dmma16816(k1[nb], ar, xsum);
        }
        __syncthreads();                                              // the slice may be overwritten two steps from now
    }
    // c[v]: row 16 mb + g + 8(v>>1), column nb * 8 + 2q + (v&1)
#pragma unroll
    for (unsigned nb = 0; nb < 8; nb++)
#pragma unroll
        for (unsigned v = 0; v < 4; v++) {
            const unsigned col = col0 + nb * 8 + 2 * q + (v & 1u);
            if (col < chunkGroups) ytmp[(uint64_t)(16 * mb + g + 8 * (v >> 1)) * chunkGroups + col] = make_double2(k1[nb][v] - k3[nb][v], k1[nb][v] + k2[nb][v]);
        }
}

// getSuperoperator (src/misc.hpp:58-81) on the device: S[(i d + k)][(j d + l)] = sum over Kraus operators, in their order, of
// conj(K[i][j]) * K[k][l]  (d = 2^t, S is d^2 x d^2). One thread per entry.
__global__ void __launch_bounds__(256) superoperatorKernel(const double2* __restrict__ kraus, unsigned numOps, unsigned d, double2* __restrict__ out) {
    const uint64_t D = (uint64_t)d * d, total = D * D;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = idx / D, col = idx - row * D;
        const unsigned i = (unsigned)(row / d), k = (unsigned)(row - (uint64_t)i * d), j = (unsigned)(col / d), l = (unsigned)(col - (uint64_t)j * d);
        double2 acc = make_double2(0.0, 0.0);
        for (unsigned o = 0; o < numOps; o++) {
            const double2* K = kraus + (uint64_t)o * d * d;
            const double2 a = K[(uint64_t)i * d + j], b = K[(uint64_t)k * d + l];
            acc = cadd(acc, make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x));
        }
        out[idx] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------

struct Gate4 { double2 m[16]; };

namespace {

// tile bits = targets U the f lowest non-target bits; localPos.pos[b] = role of the b-th (sorted) tile bit:
// i < t -> gate bit i (targets[i]), else free bit (role - t)
int buildTile(const uint32_t* targets, unsigned t, unsigned L, uint64_t targMask, unsigned f, BitSpec* tileSpec, BitSpec* localPos) {
    std::vector<uint32_t> tileBits(targets, targets + t), freeBits;
    for (unsigned b = 0; b < L && freeBits.size() < f; b++)
        if (!((targMask >> b) & 1ULL)) freeBits.push_back(b);
    tileBits.insert(tileBits.end(), freeBits.begin(), freeBits.end());
    DFSA_TRY(sortedSpec(tileBits.data(), t + f, L, tileSpec, nullptr));
    localPos->n = t + f;
    for (unsigned b = 0; b < t + f; b++) {
        unsigned q = tileSpec->pos[b], role = 0;
        bool isTarget = false;
        for (unsigned i = 0; i < t; i++) if (targets[i] == q) { role = i; isTarget = true; }
        if (!isTarget) for (unsigned i = 0; i < f; i++) if (freeBits[i] == q) role = t + i;
        localPos->pos[b] = (uint8_t)role;
    }
    return DFSA_OK;
}

// Slab layout of manyTargSpecKernel for one target placement (see the kernel's comment): per tile bit p its byte-offset
// contribution bitOff[p] = (16 << p) ^ (m[p] << 4); m[p] = 0 for p < 3, and for p >= 3 chosen so that both fragment access
// patterns are bank-conflict free: {row bit 0, row bit 1, col bit 0} and {col bit 1, col bit 2, row bit 0} must each map to
// three linearly independent bank-group vectors over GF(2) (bank group = offset bits 4..6).
template <int T>
void chooseLayout(const BitSpec& localPos, SpecLayout* z, unsigned bitOff[9]) {
    constexpr unsigned NB = SpecGeom<T>::TILE_BITS;
    unsigned posOfRow[6] = {0, 0, 0, 0, 0, 0}, posOfCol[6] = {0, 0, 0, 0, 0, 0};
    for (unsigned p = 0; p < NB; p++) {
        const unsigned role = localPos.pos[p];
        if (role < (unsigned)T) posOfRow[role] = p; else posOfCol[role - T] = p;
    }
    auto independent = [](unsigned a, unsigned b, unsigned c) { return a && b && c && a != b && a != c && b != c && (a ^ b) != c; };
    const unsigned involved[5] = {posOfRow[0], posOfRow[1], posOfCol[0], posOfCol[1], posOfCol[2]};
    unsigned m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, freePos[5], numFree = 0;
    for (unsigned k = 0; k < 5; k++) if (involved[k] >= 3) freePos[numFree++] = involved[k];
    auto vec = [&](unsigned p) { return p < 3 ? (1u << p) : m[p]; };
    unsigned combos = 1;
    for (unsigned k = 0; k < numFree; k++) combos *= 7;
    for (unsigned code = 0; code < combos; code++) {
        unsigned c = code;
        for (unsigned k = 0; k < numFree; k++) { m[freePos[k]] = 1 + c % 7; c /= 7; }
        if (independent(vec(involved[0]), vec(involved[1]), vec(involved[2])) && independent(vec(involved[3]), vec(involved[4]), vec(involved[0]))) break;
    }
    for (unsigned p = 0; p < NB; p++) bitOff[p] = (16u << p) ^ ((p >= 3 ? m[p] : 0u) << 4);
    for (unsigned i = 0; i < 6; i++) z->rowBit[i] = i < (unsigned)T ? bitOff[posOfRow[i]] : 0u;
    for (unsigned j = 0; j < 6; j++) z->colBit[j] = j < SpecGeom<T>::F ? bitOff[posOfCol[j]] : 0u;
}

// full tiles only (9 local bits): shards too small for one are the caller's business
template <int T, bool SMEM_A>
int launchSpecKernel(dfsa_state* s, const uint32_t* targets, uint64_t targMask, const double2* devGate) {
    using G = SpecGeom<T>;
    DfsaContext& ctx = dfsaCtx();
    BitSpec tileSpec, localPos;
    DFSA_TRY(buildTile(targets, T, s->logNumAmps, targMask, G::F, &tileSpec, &localPos));
    SpecLayout lay;
    unsigned bitOff[9];
    chooseLayout<T>(localPos, &lay, bitOff);
    SpecMap map;                                                     // i part of element (lane | i << 5): tile bits 5..8
    for (unsigned i = 0; i < 16; i++) {
        uint64_t off = 0;
        unsigned sOff = 0;
        for (unsigned b = 5; b < G::TILE_BITS; b++)
            if ((i >> (b - 5)) & 1u) { off |= 1ULL << tileSpec.pos[b]; sOff ^= bitOff[b]; }
        map.gByte[i] = off << 4;
        map.sByte[i] = sOff;
    }
    static bool configured = false;
    if (!configured) {
        DFSA_CUDA(cudaFuncSetAttribute(manyTargSpecKernel<T, SMEM_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smemBytes(SMEM_A)));
        configured = true;
    }
    const uint64_t numTiles = s->numAmps >> G::TILE_BITS;
    const uint64_t blocksNeeded = (numTiles + G::STREAMS - 1) / G::STREAMS;
    const unsigned grid = (unsigned)std::min<uint64_t>(blocksNeeded, (uint64_t)ctx.numSMs);
    manyTargSpecKernel<T, SMEM_A><<<grid, G::THREADS, G::smemBytes(SMEM_A), ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, localPos, devGate, map, lay);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

}  // namespace

// Host-only view of the tensor-core kernel's tile plan (no device needed): which 9 index bits form a tile, their roles
// (< t: gate-row bit, else vector bit role - t) and the shared-memory slab layout chosen for this placement.
extern "C" int dfsa_plan_manyTargLayout(const uint32_t* targets, unsigned numTargets, unsigned logNumAmps, uint32_t tileBits[9],
                                        uint32_t roles[9], uint32_t bitOff[9], uint32_t rowBit[6], uint32_t colBit[6]) {
    DFSA_REQUIRE(targets && tileBits && roles && bitOff && rowBit && colBit, "null argument");
    DFSA_REQUIRE(numTargets >= 3 && numTargets <= 6 && logNumAmps >= SpecGeom<3>::TILE_BITS && logNumAmps <= DFSA_MAX_QUBITS,
                 "the tile kernel serves 3..6 targets on shards of at least 512 amplitudes");
    BitSpec sortedT, tileSpec, localPos; uint64_t targMask;
    DFSA_TRY(sortedSpec(targets, numTargets, logNumAmps, &sortedT, &targMask));
    DFSA_TRY(buildTile(targets, numTargets, logNumAmps, targMask, SpecGeom<3>::TILE_BITS - numTargets, &tileSpec, &localPos));
    SpecLayout lay;
    unsigned off[9];
    switch (numTargets) {
        case 3:  chooseLayout<3>(localPos, &lay, off); break;
        case 4:  chooseLayout<4>(localPos, &lay, off); break;
        case 5:  chooseLayout<5>(localPos, &lay, off); break;
        default: chooseLayout<6>(localPos, &lay, off); break;
    }
    for (unsigned p = 0; p < 9; p++) { tileBits[p] = tileSpec.pos[p]; roles[p] = localPos.pos[p]; bitOff[p] = off[p]; }
    for (unsigned i = 0; i < 6; i++) { rowBit[i] = lay.rowBit[i]; colBit[i] = lay.colBit[i]; }
    return DFSA_OK;
}

// `gate` = host matrix, or -- krausOps != nullptr -- the gate is the superoperator of numKraus host Kraus operators on
// numTargets / 2 qubits, built on the device (GEMM path only; smaller superoperators are built by the caller on the host)
static int manyTargImpl(dfsa_state* s, const uint32_t* targets, unsigned numTargets, const double* gate, const double* krausOps, unsigned numKraus) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && targets && (gate || krausOps), "null argument");
    const unsigned t = numTargets, L = s->logNumAmps;
    DFSA_REQUIRE(t >= 1 && t <= L, "manyTargGate needs 1 <= numTargets <= local bits (distributed_statevector.hpp:191)");
    BitSpec sortedT; uint64_t targMask;
    DFSA_TRY(sortedSpec(targets, t, L, &sortedT, &targMask));
    const uint64_t d = 1ULL << t;
    const size_t gateBytes = d * d * sizeof(double2);
    DfsaContext& ctx = dfsaCtx();

    if (t == 1) return dfsa_k_ctrlOneTarg(s, nullptr, 0, targets[0], gate);      // same operator as K1: pair-stream kernel

    if (t == 2) {
        // register-resident 4x4 complex matvec on the streaming skeleton: item = the four amplitudes spanned by the two
        // target bits (gate bit 0 <-> targets[0], bit 1 <-> targets[1]); the 16 gate entries ride in the kernel parameters
        Gate4 g;
        for (int e = 0; e < 16; e++) g.m[e] = hostAmp(gate + 2 * e);
        double2* amps = s->arr[DFSA_AMPS];
        const uint64_t b0 = 1ULL << targets[0], b1 = 1ULL << targets[1];
        const unsigned lo = sortedT.pos[0], hi = sortedT.pos[1];
        auto ld = [=] __device__(uint64_t j) {
            QuadAt q;
            q.idx = insertZeroBit(insertZeroBit(j, lo), hi);
            q.v0 = amps[q.idx]; q.v1 = amps[q.idx | b0]; q.v2 = amps[q.idx | b1]; q.v3 = amps[q.idx | b0 | b1];
            return q;
        };
        auto st = [=] __device__(uint64_t, const QuadAt& q) {
            const double2 x[4] = {q.v0, q.v1, q.v2, q.v3};
            double2 y[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                double2 acc = cmul(g.m[4 * r], x[0]);
#pragma unroll
                for (int l = 1; l < 4; l++) acc = cfma(g.m[4 * r + l], x[l], acc);
                y[r] = acc;
            }
            amps[q.idx] = y[0]; amps[q.idx | b0] = y[1]; amps[q.idx | b1] = y[2]; amps[q.idx | b0 | b1] = y[3];
        };
        return launchStream<1, QuadAt>(s->numAmps >> 2, ld, st);
    }

    // tensor-core tiles span 9 local bits; a smaller shard (< 512 amplitudes) goes to the generic kernel below
    if (t >= 3 && t <= 6 && L >= SpecGeom<3>::TILE_BITS) {
        void* stage; int slot;
        DFSA_TRY(dfsaStagingAcquire(gateBytes, &stage, &slot));
        memcpy(stage, gate, gateBytes);
        double2* dev;
        DFSA_TRY(dfsaScratch(gateBytes, &dev));
        DFSA_CUDA(cudaMemcpyAsync(dev, stage, gateBytes, cudaMemcpyHostToDevice, ctx.compute));
        DFSA_TRY(dfsaStagingCommit(slot));
        switch (t) {
            case 3:  return launchSpecKernel<3, false>(s, targets, targMask, dev);
            case 4:  return launchSpecKernel<4, false>(s, targets, targMask, dev);
            // t = 5 measured both ways at 30 qubits: A-fragments in registers 7.2 ms, in shared memory 7.5 ms
            case 5:  return launchSpecKernel<5, false>(s, targets, targMask, dev);
            default: return launchSpecKernel<6, true>(s, targets, targMask, dev);
        }
    }

    DFSA_REQUIRE(gate || (t >= 7 && t <= 11), "device-built superoperators are served by the GEMM path (7..11 effective targets)");
    if (t >= 7 && t <= 11 && (!gate || !getenv("DFSA_MANYTARG_GENERIC"))) {
        // scratch: [gate D^2 amps][A-fragments 3 D^2 doubles][row offsets D u64][result slab]
        const size_t fragBytes = 3 * d * d * sizeof(double), offBytes = d * sizeof(uint64_t);
        const uint64_t numGroups = s->numAmps >> t;
        uint64_t chunkGroups = std::min<uint64_t>(numGroups, std::max<uint64_t>(GEMM_BN, ((256ull << 20) / sizeof(double2)) >> t));
        chunkGroups = std::min<uint64_t>(chunkGroups, 0x7fffffc0ull);
        // (the result slab doubles as the landing area of the uploaded Kraus operators, so it is at least that large)
        const size_t krausBytes = krausOps ? (size_t)numKraus * (d * sizeof(double2)) : 0;      // numKraus matrices of sqrt(d) x sqrt(d)
        const size_t slabBytes = std::max((size_t)chunkGroups * d * sizeof(double2), krausBytes);
        double2* dev;
        DFSA_TRY(dfsaScratch(gateBytes + fragBytes + offBytes + slabBytes + 1024, &dev));
        double2* dGate = dev;
        double2* dFrag = (double2*)((char*)dev + gateBytes);
        uint64_t* dOff = (uint64_t*)((char*)dFrag + fragBytes);
        double2* dSlab = (double2*)((char*)dOff + ((offBytes + 255) / 256) * 256);
        std::vector<uint64_t> rowOff(d);
        for (uint64_t r = 0; r < d; r++) {
            uint64_t off = 0;
            for (unsigned b = 0; b < t; b++) off |= ((r >> b) & 1ULL) << targets[b];
            rowOff[r] = off;
        }
        if (gate) DFSA_CUDA(cudaMemcpyAsync(dGate, gate, gateBytes, cudaMemcpyHostToDevice, ctx.compute));
        else {
            // the Kraus operators (numKraus x 2^(t/2) x 2^(t/2)) ride in the result slab, which is not in use yet
            const unsigned dk = 1u << (t / 2);
            DFSA_CUDA(cudaMemcpyAsync(dSlab, krausOps, krausBytes, cudaMemcpyHostToDevice, ctx.compute));
            superoperatorKernel<<<ctx.numSMs * 4, 256, 0, ctx.compute>>>(dSlab, numKraus, dk, dGate);
            DFSA_LAUNCH_CHECK();
        }
        DFSA_CUDA(cudaMemcpyAsync(dOff, rowOff.data(), offBytes, cudaMemcpyHostToDevice, ctx.compute));
        DFSA_CUDA(cudaStreamSynchronize(ctx.compute));        // caller-owned pageable sources
        gemmPrepKernel<<<ctx.numSMs * 4, 256, 0, ctx.compute>>>(dGate, (unsigned)d, dFrag);
        DFSA_LAUNCH_CHECK();
        double2* amps = s->arr[DFSA_AMPS];
        for (uint64_t g0 = 0; g0 < numGroups; g0 += chunkGroups) {
            const unsigned cg = (unsigned)std::min<uint64_t>(chunkGroups, numGroups - g0);
            dim3 grid((cg + GEMM_BN - 1) / GEMM_BN, (unsigned)(d / GEMM_BM));
            if (targets[0] == 0) manyTargGemmKernel<true><<<grid, GEMM_THREADS, 0, ctx.compute>>>(amps, dSlab, g0, cg, sortedT, dOff, (unsigned)d, dFrag);
            else manyTargGemmKernel<false><<<grid, GEMM_THREADS, 0, ctx.compute>>>(amps, dSlab, g0, cg, sortedT, dOff, (unsigned)d, dFrag);
            DFSA_LAUNCH_CHECK();
            // results home: amps[group | row bits] = slab[row][group]
            const uint64_t* offs = dOff;
            const BitSpec sp = sortedT;
            auto ld = [=] __device__(uint64_t i) { return Amp1{dSlab[i]}; };
            auto st = [=] __device__(uint64_t i, const Amp1& v) {
                const uint64_t r = i / cg, gc = i - r * cg;
                amps[insertZeroBits(g0 + gc, sp) | offs[r]] = v.a;
            };
            DFSA_TRY((launchStream<2, Amp1>((uint64_t)cg * d, ld, st)));
        }
        return DFSA_OK;
    }

    DFSA_REQUIRE(d * sizeof(double2) <= 200 * 1024, "manyTargGate: 2^numTargets amplitudes must fit shared memory (numTargets <= 13)");
    BitSpec caller; caller.n = t;
    for (unsigned i = 0; i < t; i++) caller.pos[i] = (uint8_t)targets[i];
    const uint64_t numGroups = s->numAmps >> t;
    const unsigned grid = (unsigned)std::min<uint64_t>(numGroups, (uint64_t)ctx.numSMs * 2);
    double2* dev;
    DFSA_TRY(dfsaScratch(gateBytes + (size_t)grid * d * sizeof(double2), &dev));
    DFSA_CUDA(cudaMemcpyAsync(dev, gate, gateBytes, cudaMemcpyHostToDevice, ctx.compute));
    DFSA_CUDA(cudaStreamSynchronize(ctx.compute));        // caller-owned pageable source of unbounded size
    const size_t smemBytes = d * sizeof(double2);
    DFSA_CUDA(cudaFuncSetAttribute(manyTargGenericKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    manyTargGenericKernel<<<grid, 256, smemBytes, ctx.compute>>>(s->arr[DFSA_AMPS], numGroups, sortedT, caller, t, dev, dev + d * d);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

extern "C" int dfsa_k_manyTarg(dfsa_state* s, const uint32_t* targets, unsigned numTargets, const double* gate) {
    DFSA_REQUIRE(gate, "null gate");
    return manyTargImpl(s, targets, numTargets, gate, nullptr, 0);
}

// krausMap's local part (distributed_densitymatrix.hpp:79-89): the 2t-target gate sum_K conj(K) (x) K on the bits
// {targets, targets + N} (all local by now). Small superoperators (2t <= 6) are built here on the host exactly as
// getSuperoperator does (misc.hpp:58-81); from 2t = 8 on the 4^t x 4^t matrix is built ON THE DEVICE from the uploaded Kraus
// operators and consumed by the GEMM kernel without ever existing on the host.
extern "C" int dfsa_k_krausMap(dfsa_state* s, const uint32_t* targets2t, unsigned numTargets2t, const double* krausOps, unsigned numOps) {
    DFSA_REQUIRE(s && targets2t && krausOps && numOps >= 1 && numTargets2t >= 2 && (numTargets2t & 1u) == 0, "bad argument");
    if (numTargets2t >= 8 && numTargets2t <= 11) return manyTargImpl(s, targets2t, numTargets2t, nullptr, krausOps, numOps);
    const uint64_t d = 1ULL << (numTargets2t / 2), D = d * d;
    std::vector<double> super(2 * D * D, 0.0);
    for (unsigned o = 0; o < numOps; o++) {
        const double* K = krausOps + 2 * d * d * o;
        for (uint64_t i = 0; i < d; i++)
            for (uint64_t j = 0; j < d; j++) {
                const double cr = K[2 * (i * d + j)], ci = -K[2 * (i * d + j) + 1];
                for (uint64_t k = 0; k < d; k++)
                    for (uint64_t l = 0; l < d; l++) {
                        const double br = K[2 * (k * d + l)], bi = K[2 * (k * d + l) + 1];
                        double* e = &super[2 * ((i * d + k) * D + (j * d + l))];
                        e[0] += cr * br - ci * bi;
                        e[1] += cr * bi + ci * br;
                    }
            }
    }
    return manyTargImpl(s, targets2t, numTargets2t, super.data(), nullptr, 0);
}
