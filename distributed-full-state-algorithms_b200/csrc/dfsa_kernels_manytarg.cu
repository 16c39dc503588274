// dfsa_kernels_manytarg.cu -- K4: the dense 2^t x 2^t gate of local_statevector.hpp:72-99.
// Row/column bit i of the gate belongs to targets[i] in the CALLER's order; amplitudes outside the targets are
// enumerated by inserting zeros at the sorted target positions. 32*A bytes of HBM traffic and 8*2^t flop per
// amplitude: HBM-bound for t <= 4, FP64-pipe-bound from t = 5 (256 flop per 32 bytes vs a ridge of ~5.7 flop/B).
//
// Kernels:
//   t == 1  the pair-stream kernel of K1 (same operator)
//   t == 2  register-resident 4x4 matvec on the streaming skeleton (quad items, gate in kernel parameters); HBM-bound
//   t = 3..5 manyTargDmmaKernel<T>  warp-private tiles of 2^t x 32 amplitudes, FP64 tensor cores
//           (mma.sync m16n8k16.f64 -> SASS DMMA.8x8x4), gate held as A-fragments in registers, cp.async double buffering.
//           ncu showed the DFMA version of t=5 to be issue/I-cache bound at 18 % FP64-pipe utilisation and 8 % of DRAM
//           bandwidth, i.e. compute- not HBM-bound (profiles/r01_ncu_manytarg.txt): the case north_star reserves the
//           tensor path for. t=3,4 use the same pipeline and are HBM-bound.
//   t == 6  manyTargTileKernel<8>   block tile, gate transposed in shared memory, DFMA
//   t >= 7  manyTargGenericKernel   one block per 2^t group, gate streamed from L2, warp-per-row reduction
// A tile = the 2^(t+f) amplitudes spanned by the t target bits and the f (<= 5) lowest non-target bits, so global
// traffic is contiguous runs and each of the 32 lanes of a warp owns one vector; it is staged as X[row][lane]
// (row = gate-ordered target bits).
#include <algorithm>
#include <vector>

#include <string.h>
#include "dfsa_stream_kernels.cuh"

// ---------------------------------------------------------------------------------------------------------
// t == 5 on the FP64 tensor cores. The complex 32x32 matvec over the 32 vectors of a tile is the real GEMM
//   [C_re; C_im] (64 x 32) = [[G_re, -G_im], [G_im, G_re]] (64 x 64) * [X_re; X_im] (64 x 32),
// issued as m16n8k16 f64 MMAs (SASS: DMMA.8x8x4 x8). One warp owns one tile; G_re and G_im live in registers as
// A-fragments for the whole kernel (2 x 2 blocks of 16x16 each: 128 registers), the tile is staged through the warp's
// private shared-memory slab X[l][n] (row stride 34 amplitudes: the 4x8 B-fragment footprint then hits 8 distinct
// 16-byte bank groups per quarter-warp), and one LDS.128 yields both the X_re and the X_im fragment element.
// Fragment maps (m16n8k16.row.col.f64; g = lane/4, q = lane%4):
//   a[v]: row g + 8(v&1),  col q + 4(v>>1)      b[v]: k = q + 4v, n = g      c[v]: row g + 8(v>>1), col 2q + (v&1)
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA-style 1-D bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion: one instruction moves a whole
// 512-byte tile row between global and shared memory through the async proxy
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bulkStore(void* gmemDst, const void* smemSrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmemDst), "r"(smemAddr(smemSrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitRead0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-T geometry. A warp owns two slabs (double buffer) of D rows x VEC vectors, row stride VEC + 2 amplitudes (the 4x8
// B-fragment footprint then hits 8 distinct 16-byte bank groups per quarter-warp). t = 5 is register-limited to 8 warps
// per SM (238 registers), so it uses 16-vector tiles: 8 x 2 x 32 x 18 x 16 B = 144 KiB lets all 8 warps (2 per
// scheduler) be resident; t = 3, 4 use 32-vector tiles and 6-warp blocks (4 resp. 2 blocks per SM).
template <int T> struct DmmaGeom {
    static constexpr unsigned F = (T == 5) ? 4 : 5;          // free (vector) bits per tile
    static constexpr unsigned VEC = 1u << F;
    static constexpr unsigned S = VEC + 2;
    static constexpr unsigned WARPS = (T == 5) ? 8 : 6;
    static constexpr unsigned D = 1u << T;
    static constexpr unsigned EPL = D * VEC / 32;             // tile elements each lane moves (8 or 16)
    static constexpr size_t smemBytes = (size_t)WARPS * 2 * D * S * sizeof(double2);
};

// Where element (lane | i << 5) of a tile lives, split into its lane part (registers) and its i part (this table). The
// table rides in the kernel parameters, so after unrolling every entry is a constant-bank operand of the address add:
// one cp.async / st.global costs three integer instructions instead of two shared-memory table reads plus ~15 ALU ops
// (ncu, round 1: 1200 non-DMMA instructions per t=5 tile against 512 DMMA slots kept the tensor pipe at 54 %).
template <int T> struct TileMap {
    uint64_t gByte[DmmaGeom<T>::EPL];                         // byte offset in the shard
    uint32_t sByte[DmmaGeom<T>::EPL];                         // byte offset in the slab X[row][n]
};

// Software pipeline per warp: while the tensor cores work on tile k (slab k&1), cp.async (LDGSTS, L2 -> shared,
// no registers) is already filling the other slab with tile k+1, so HBM latency is hidden with few resident warps.
//   T = 5: complex blocks, [mb][kb] = 2 x 2 blocks of 16x16 for G_re and G_im (128 registers), 16 MMAs per 8 vectors
//   T = 4: one 16x16 block each for G_re, G_im, 4 MMAs per 8 vectors
//   T = 3: the 8x8 complex gate as ONE real 16x16 A-fragment [[G_re,-G_im],[G_im,G_re]], B = [X_re; X_im] stacked
//          along k, 1 MMA per 8 vectors
//   BULK: the tile's 32 vectors are contiguous in memory (all targets >= bit 5, f == 5), so each of the 2^T tile rows is
//         one 512-byte run: lane i moves row i with a single bulk async copy in (mbarrier completion) and out
//         (bulk async-group), instead of 32 per-lane 16-byte cp.async / st.global each.
template <int T, bool BULK>
__global__ void __launch_bounds__(32 * DmmaGeom<T>::WARPS, (T == 5) ? 1 : (T == 4 ? 2 : 4))
manyTargDmmaKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec, BitSpec localPos, const double2* __restrict__ gate, TileMap<T> map) {
    constexpr unsigned D = 1u << T, S = DmmaGeom<T>::S, VEC = DmmaGeom<T>::VEC, WARPS = DmmaGeom<T>::WARPS;
    constexpr unsigned F = DmmaGeom<T>::F, EPL = DmmaGeom<T>::EPL, SLAB_BYTES = D * S * 16u;
    constexpr int NB = (T >= 4) ? (int)(D / 16) : 1;              // 16-row blocks of the complex gate (T >= 4)
    extern __shared__ double2 smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned g = lane >> 2, q = lane & 3u;
    double2* slab = smem + (size_t)warp * (2 * D * S);

    // tile element e (bit b of e = b-th lowest tile bit): shard offset, slab row (gate-ordered target bits) and slab column
    auto decompose = [&](unsigned e, uint64_t& off, unsigned& row, unsigned& n) {
        off = 0; row = 0; n = 0;
        for (unsigned b = 0; b < T + F; b++) {
            const unsigned bit = (e >> b) & 1u, role = localPos.pos[b];
            off |= (uint64_t)bit << tileSpec.pos[b];
            if (role < (unsigned)T) row |= bit << role; else n |= bit << (role - T);
        }
    };

    // the gate as A-fragments (a[v]: row g + 8(v&1), col q + 4(v>>1))
    double gr[NB][NB][8], gi[NB][NB][8];
    double gd[8], gs[8];                                           // T == 4 (3M form): G_im - G_re and G_re + G_im
    if constexpr (T >= 4) {
#pragma unroll
        for (int mb = 0; mb < NB; mb++)
#pragma unroll
            for (int kb = 0; kb < NB; kb++)
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    const double2 e = gate[(16 * mb + g + 8 * (v & 1)) * D + 16 * kb + q + 4 * (v >> 1)];
                    gr[mb][kb][v] = e.x;
                    gi[mb][kb][v] = e.y;
                }
        if constexpr (T == 4) {
#pragma unroll
            for (int v = 0; v < 8; v++) { gd[v] = gi[0][0][v] - gr[0][0][v]; gs[v] = gr[0][0][v] + gi[0][0][v]; }
        }
    } else {
        // T == 3: real 16x16 matrix R = [[G_re, -G_im], [G_im, G_re]] in gr[0][0]; gi unused
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const unsigned row = g + 8 * (v & 1), col = q + 4 * (v >> 1);
            const double2 e = gate[(row & 7u) * D + (col & 7u)];
            gr[0][0][v] = ((row < 8) == (col < 8)) ? e.x : ((row < 8) ? -e.y : e.y);
            gi[0][0][v] = 0.0;
        }
    }

    const uint64_t stride = (uint64_t)gridDim.x * WARPS;
    __shared__ uint64_t bars[WARPS][2];                            // BULK: one mbarrier per warp and slab
    __shared__ uint64_t rowOff[D];                                 // BULK: shard offset of tile row r (element r << F)
    __shared__ unsigned rowSlab[D];                                //       and the slab row it lands in
    unsigned phase[2] = {0u, 0u};
    // generic path: this lane's share of every element address (the i part comes from `map`)
    char* laneG = reinterpret_cast<char*>(amps);
    unsigned laneS = smemAddr(slab);
    if constexpr (BULK) {
        if (threadIdx.x < D) {
            uint64_t off; unsigned row, n;
            decompose(threadIdx.x << F, off, row, n);
            rowOff[threadIdx.x] = off; rowSlab[threadIdx.x] = row;
        }
        __syncthreads();
        if (lane == 0) { mbarInit(&bars[warp][0], 1); mbarInit(&bars[warp][1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    } else {
        uint64_t off; unsigned row, n;
        decompose(lane, off, row, n);
        laneG += off << 4;
        laneS += (row * S + n) << 4;
    }
    // row i of a tile (BULK): global run at base | rowOff[i], slab row rowSlab[i]
    auto prefetch = [&](uint64_t base, unsigned b) {
        if constexpr (BULK) {
            double2* X = slab + (size_t)b * (D * S);
            if (lane == 0) mbarExpectTx(&bars[warp][b], D * VEC * 16u);
            if (lane < D) bulkLoad(&X[rowSlab[lane] * S], &amps[base | rowOff[lane]], VEC * 16u, &bars[warp][b]);
        } else {
            const char* src = laneG + (base << 4);
            const unsigned dst = laneS + b * SLAB_BYTES;
#pragma unroll
            for (unsigned i = 0; i < EPL; i++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + map.sByte[i]), "l"(src + map.gByte[i]) : "memory");
            cpAsyncCommit();
        }
    };

    // shard offset of a tile's element 0: the tile counter with zeros inserted at the T + F tile bits (positions are
    // constant-bank operands after unrolling); computed once per tile, when it is prefetched, and carried to its store
    uint64_t tile = (uint64_t)blockIdx.x * WARPS + warp;
    uint64_t base = 0, nextBase = insertZeroBitsN<T + F>(tile, tileSpec);
    if (tile < numTiles) prefetch(nextBase, 0);
    for (unsigned it = 0; tile < numTiles; it++, tile += stride) {
        const unsigned cur = it & 1u;
        double2* X = slab + (size_t)cur * (D * S);
        const bool more = tile + stride < numTiles;
        base = nextBase;
        nextBase = insertZeroBitsN<T + F>(tile + stride, tileSpec);
        if constexpr (BULK) {
            // the other slab was the source of the previous tile's bulk stores: they must have finished reading it
            bulkWaitRead0();
            __syncwarp();
            if (more) prefetch(nextBase, cur ^ 1u);
            mbarWait(&bars[warp][cur], phase[cur]);
            phase[cur] ^= 1u;
        } else {
            if (more) { prefetch(nextBase, cur ^ 1u); cpAsyncWait<1>(); }
            else cpAsyncWait<0>();
            __syncwarp();
        }
#pragma unroll 1
        for (unsigned nb = 0; nb < VEC / 8; nb++) {
            if constexpr (T == 4) {
                // 3M complex product: k1 = G_re (X_re + X_im), k2 = (G_im - G_re) X_re, k3 = (G_re + G_im) X_im;
                // re = k1 - k3, im = k1 + k2 -- three real MMAs per 8 vectors instead of four
                double k1[4] = {0.0, 0.0, 0.0, 0.0}, k2[4] = {0.0, 0.0, 0.0, 0.0}, k3[4] = {0.0, 0.0, 0.0, 0.0};
                double xr[4], xi[4], xs[4];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const double2 x = X[(q + 4 * v) * S + nb * 8 + g];
                    xr[v] = x.x; xi[v] = x.y; xs[v] = x.x + x.y;
                }
                dmma16816(k1, gr[0][0], xs);
                dmma16816(k2, gd, xr);
                dmma16816(k3, gs, xi);
                __syncwarp();                                       // every lane has read this n-block's columns
#pragma unroll
                for (int v = 0; v < 4; v++)
                    X[(g + 8 * (v >> 1)) * S + nb * 8 + 2 * q + (v & 1)] = make_double2(k1[v] - k3[v], k1[v] + k2[v]);
            } else if constexpr (T == 5) {
                double cre[NB][4], cim[NB][4];
#pragma unroll
                for (int mb = 0; mb < NB; mb++)
#pragma unroll
                    for (int v = 0; v < 4; v++) { cre[mb][v] = 0.0; cim[mb][v] = 0.0; }
#pragma unroll
                for (int kb = 0; kb < NB; kb++) {
                    double xr[4], xi[4], nxi[4];
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const double2 x = X[(16 * kb + q + 4 * v) * S + nb * 8 + g];
                        xr[v] = x.x; xi[v] = x.y; nxi[v] = -x.y;
                    }
                    // independent accumulator chains issued round-robin: consecutive MMAs never depend on each other
#pragma unroll
                    for (int mb = 0; mb < NB; mb++) {
                        dmma16816(cre[mb], gr[mb][kb], xr);
                        dmma16816(cim[mb], gi[mb][kb], xr);
                    }
#pragma unroll
                    for (int mb = 0; mb < NB; mb++) {
                        dmma16816(cre[mb], gi[mb][kb], nxi);
                        dmma16816(cim[mb], gr[mb][kb], xi);
                    }
                }
                __syncwarp();                                       // every lane has read this n-block's columns
#pragma unroll
                for (int mb = 0; mb < NB; mb++)
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        X[(16 * mb + g + 8 * (v >> 1)) * S + nb * 8 + 2 * q + (v & 1)] = make_double2(cre[mb][v], cim[mb][v]);
            } else {
                // b[v]: k = q + 4v, n = g ; k < 8 -> X_re row k, k >= 8 -> X_im row k-8
                const double2 x0 = X[q * S + nb * 8 + g], x1 = X[(q + 4) * S + nb * 8 + g];
                const double bfrag[4] = {x0.x, x1.x, x0.y, x1.y};
                double c[4] = {0.0, 0.0, 0.0, 0.0};
                dmma16816(c, gr[0][0], bfrag);
                __syncwarp();
                // c[v]: row g + 8(v>>1) (rows 8..15 = imaginary parts of complex row g), col 2q + (v&1)
                X[g * S + nb * 8 + 2 * q] = make_double2(c[0], c[2]);
                X[g * S + nb * 8 + 2 * q + 1] = make_double2(c[1], c[3]);
            }
        }
        if constexpr (BULK) {
            fenceProxyAsync();                                      // results written with st.shared -> visible to the async proxy
            __syncwarp();
            if (lane < D) bulkStore(&amps[base | rowOff[lane]], &X[rowSlab[lane] * S], VEC * 16u);
            bulkCommit();
        } else {
            __syncwarp();
            char* dst = laneG + (base << 4);
            const unsigned src = laneS + cur * SLAB_BYTES;
#pragma unroll
            for (unsigned i0 = 0; i0 < EPL; i0 += 4) {                // four shared loads in flight, then their four stores
                double2 v[4];
#pragma unroll
                for (unsigned i = 0; i < 4; i++)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[i].x), "=d"(v[i].y) : "r"(src + map.sByte[i0 + i]) : "memory");
#pragma unroll
                for (unsigned i = 0; i < 4; i++) *reinterpret_cast<double2*>(dst + map.gByte[i0 + i]) = v[i];
            }
            __syncwarp();
        }
    }
    if constexpr (BULK) bulkWaitRead0();
}

// ---------------------------------------------------------------------------------------------------------
// t == 5, second form: 3M complex product, gate rows split across a WARP PAIR.
// The one-warp kernel above needs all of G_re and G_im as A-fragments (128 registers), which caps the SM at 8 warps and
// rules out a third matrix. Here two warps share a tile and each owns 16 of the 32 gate rows, so the three matrices of the
// 3M product -- G_re, G_im - G_re, G_re + G_im -- cost 96 registers, 12 warps fit, and the tensor pipe does 3 real MMAs
// where the 4M form does 4 (FP64 bound 192 instead of 256 flop per amplitude, about level with the HBM bound):
//   k1 = G_re (X_re + X_im), k2 = (G_im - G_re) X_re, k3 = (G_re + G_im) X_im;  Y_re = k1 - k3, Y_im = k1 + k2.
// Per pair: two input slabs (cp.async double buffer; each warp fetches half of the tile) and one output slab, so results
// never overwrite operands and two named barriers per tile suffice: A = "tile landed, output slab free",
// B = "output slab complete". Element (lane | h << 5 | j << 6) of a tile, h = warp of the pair: the lane and h parts of
// its addresses are per-thread registers, the j part is the table in the kernel parameters.
struct Pair5 {
    static constexpr unsigned T = 5, D = 32, F = 4, VEC = 16, S = VEC + 2, PAIRS = 6, EPW = D * VEC / 64;   // 8 elements per thread
    static constexpr unsigned SLAB = D * S, SLAB_BYTES = SLAB * 16u;
    static constexpr size_t smemBytes = (size_t)PAIRS * 3 * SLAB_BYTES;       // 162 KiB
};
struct Pair5Map { uint64_t gByte[Pair5::EPW]; uint32_t sByte[Pair5::EPW]; };

__global__ void __launch_bounds__(64 * Pair5::PAIRS, 1)
manyTarg5PairKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec, BitSpec localPos, const double2* __restrict__ gate, Pair5Map map) {
    constexpr unsigned T = Pair5::T, D = Pair5::D, F = Pair5::F, S = Pair5::S, VEC = Pair5::VEC, EPW = Pair5::EPW;
    constexpr unsigned SLAB = Pair5::SLAB, SLAB_BYTES = Pair5::SLAB_BYTES;
    extern __shared__ double2 smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, pair = warp >> 1, h = warp & 1u;
    const unsigned g = lane >> 2, q = lane & 3u;
    double2* in0 = smem + (size_t)pair * (3 * SLAB);
    double2* out = in0 + 2 * SLAB;

    // this thread's share of every element address: tile bits 0..4 = lane, bit 5 = h
    uint64_t laneOff = 0;
    unsigned laneRow = 0, laneN = 0;
    {
        const unsigned e = lane | (h << 5);
#pragma unroll
        for (unsigned b = 0; b < 6; b++) {
            const unsigned bit = (e >> b) & 1u, role = localPos.pos[b];
            laneOff |= (uint64_t)bit << tileSpec.pos[b];
            if (role < T) laneRow |= bit << role; else laneN |= bit << (role - T);
        }
    }
    char* laneG = reinterpret_cast<char*>(amps) + (laneOff << 4);
    const unsigned laneS = (laneRow * S + laneN) << 4;
    const unsigned inS = smemAddr(in0) + laneS, outS = smemAddr(out) + laneS;

    // A-fragments of this warp's 16 gate rows (a[v]: row g + 8(v&1), col q + 4(v>>1)), k-blocks 0 and 1
    double ar[2][8], ad[2][8], as[2][8];
#pragma unroll
    for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const double2 e = gate[(16 * h + g + 8 * (v & 1)) * D + 16 * kb + q + 4 * (v >> 1)];
            ar[kb][v] = e.x;
            ad[kb][v] = e.y - e.x;
            as[kb][v] = e.x + e.y;
        }

    auto prefetch = [&](uint64_t base, unsigned b) {
        const char* src = laneG + (base << 4);
        const unsigned dst = inS + b * SLAB_BYTES;
#pragma unroll
        for (unsigned j = 0; j < EPW; j++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + map.sByte[j]), "l"(src + map.gByte[j]) : "memory");
        cpAsyncCommit();
    };
    auto pairBarrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1u) : "memory"); };

    const uint64_t stride = (uint64_t)gridDim.x * Pair5::PAIRS;
    uint64_t tile = (uint64_t)blockIdx.x * Pair5::PAIRS + pair;
    uint64_t base = 0, nextBase = insertZeroBitsN<T + F>(tile, tileSpec);
    if (tile < numTiles) prefetch(nextBase, 0);
    for (unsigned it = 0; tile < numTiles; it++, tile += stride) {
        const unsigned cur = it & 1u;
        const double2* X = in0 + (size_t)cur * SLAB;
        base = nextBase;
        nextBase = insertZeroBitsN<T + F>(tile + stride, tileSpec);
        // the other input slab was last read before barrier B of the previous tile, which this warp has passed
        if (tile + stride < numTiles) { prefetch(nextBase, cur ^ 1u); cpAsyncWait<1>(); }
        else cpAsyncWait<0>();
        pairBarrier();                                              // A: both halves of this tile landed; output slab drained
#pragma unroll 1
        for (unsigned nb = 0; nb < VEC / 8; nb++) {
            double k1[4] = {0.0, 0.0, 0.0, 0.0}, k2[4] = {0.0, 0.0, 0.0, 0.0}, k3[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int kb = 0; kb < 2; kb++) {
                double xr[4], xi[4], xs[4];
#pragma unroll
                for (int v = 0; v < 4; v++) {                       // b[v]: k = q + 4v, n = g
                    const double2 x = X[(16 * kb + q + 4 * v) * S + nb * 8 + g];
                    xr[v] = x.x; xi[v] = x.y; xs[v] = x.x + x.y;
                }
                dmma16816(k1, ar[kb], xs);
                dmma16816(k2, ad[kb], xr);
                dmma16816(k3, as[kb], xi);
            }
#pragma unroll
            for (int v = 0; v < 4; v++)                             // c[v]: row g + 8(v>>1), col 2q + (v&1)
                out[(16 * h + g + 8 * (v >> 1)) * S + nb * 8 + 2 * q + (v & 1)] = make_double2(k1[v] - k3[v], k1[v] + k2[v]);
        }
        pairBarrier();                                              // B: all 32 rows of the output slab are written
        char* dst = laneG + (base << 4);
#pragma unroll
        for (unsigned j0 = 0; j0 < EPW; j0 += 4) {
            double2 v[4];
#pragma unroll
            for (unsigned j = 0; j < 4; j++)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[j].x), "=d"(v[j].y) : "r"(outS + map.sByte[j0 + j]) : "memory");
#pragma unroll
            for (unsigned j = 0; j < 4; j++) *reinterpret_cast<double2*>(dst + map.gByte[j0 + j]) = v[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// t == 5, third form: the pair kernel's arithmetic with the memory traffic moved to dedicated MOVER warps.
// ncu on the pair kernel: tensor pipe 70 % active; the three warps of a scheduler share the pipe fairly, so they drift
// into lock-step -- all in their DMMA phase together, then all in their load/store phase together with the pipe idle.
// Here 8 compute warps (4 pairs, gate rows split as above) run nothing but LDS -> DMMA -> STS, and 4 mover warps (one
// per pair, one per scheduler) do every cp.async, shared-memory drain and global store. Per pair: NIN input slabs
// and two output slabs, handed over through mbarriers in shared memory:
//   full[s]    mover -> compute: tile landed in input slab s (cp.async.mbarrier.arrive.noinc of the 32 mover lanes)
//   done[o]    compute -> mover: output slab o holds the results of a tile and its input slab is free again
//   drained[o] mover -> compute: output slab o has been read out, it may be overwritten
// Tile i of a pair uses input slab i % NIN and output slab i % 2; the mover loads tile i + NIN as soon as tile i is
// done, i.e. NIN - 1 tile-times ahead of its use.
//
// Slab layout: X[row][col], 16 columns of 16 bytes per row, NO padding; instead col = n ^ sw(row) with
// sw(row) = XOR of c[i] over the set bits i of row. c[0] = 5 and c[1] = 6 make the B-fragment reads (rows q + const,
// columns g + const) and the result writes (rows g + const, columns 2q + const) bank-conflict free; c[2..4] are chosen
// per launch so that the mover's scatter is conflict free too: its lanes follow ADDRESS order (coalesced global access),
// and which of the low address bits are gate rows depends on the targets (ncu, low targets, padded layout: 2.6x the
// L2 read sectors and 1.6x the shared wavefronts of the high-target case, all replays of conflicting cp.async).
// Slabs are 8 KiB and 8 KiB-aligned, and row / column fields of an offset never overlap, so every address is
// slab ^ (thread part) ^ (instruction part): one LOP3.
struct Spec5 {
    static constexpr unsigned T = 5, D = 32, F = 4, VEC = 16, PAIRS = 4, NOUT = 2;
    static constexpr unsigned SLAB_BYTES = D * VEC * 16u;                       // 8 KiB
    static constexpr unsigned THREADS = 32 * (2 * PAIRS + PAIRS);              // 8 compute warps + 4 mover warps
    static constexpr size_t smemBytes(unsigned nin) { return (size_t)PAIRS * (nin + NOUT) * SLAB_BYTES + SLAB_BYTES; }   // + alignment slack
};
struct Spec5Swz { unsigned c[5]; unsigned viaL1; };   // viaL1: cp.async.ca instead of .cg (see launchSpec5)

__host__ __device__ __forceinline__ unsigned spec5Sw(unsigned row, const Spec5Swz& z) {
    unsigned v = 0;
#pragma unroll
    for (unsigned i = 0; i < 5; i++) v ^= ((row >> i) & 1u) ? z.c[i] : 0u;
    return v;
}
// byte offset of (row, n) inside a slab
__host__ __device__ __forceinline__ unsigned spec5Offset(unsigned row, unsigned n, const Spec5Swz& z) {
    return (row << 8) | ((n ^ spec5Sw(row, z)) << 4);
}

__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void cpAsyncArriveNoinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}

template <unsigned NIN, bool UNROLL_NB>
__global__ void __launch_bounds__(Spec5::THREADS, 1)
manyTarg5SpecKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec, BitSpec localPos, const double2* __restrict__ gate, TileMap<5> map, Spec5Swz swz) {
    constexpr unsigned T = Spec5::T, D = Spec5::D, F = Spec5::F, VEC = Spec5::VEC, PAIRS = Spec5::PAIRS;
    constexpr unsigned SLAB_BYTES = Spec5::SLAB_BYTES, NOUT = Spec5::NOUT;
    extern __shared__ double2 smem[];
    __shared__ uint64_t full[PAIRS][NIN], done[PAIRS][NOUT], drained[PAIRS][NOUT];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool mover = warp >= 2 * PAIRS;
    const unsigned pair = mover ? warp - 2 * PAIRS : warp >> 1, h = warp & 1u;
    // shared-window addresses of this pair's slabs (8 KiB-aligned)
    const unsigned slab0 = ((smemAddr(smem) + SLAB_BYTES - 1u) & ~(SLAB_BYTES - 1u)) + pair * ((NIN + NOUT) * SLAB_BYTES);
    const unsigned in0 = slab0, out0 = slab0 + NIN * SLAB_BYTES;

    if (threadIdx.x < PAIRS) {
        for (unsigned s = 0; s < NIN; s++) mbarInit(&full[threadIdx.x][s], 32);       // the mover's 32 lanes
        for (unsigned o = 0; o < NOUT; o++) { mbarInit(&done[threadIdx.x][o], 2); mbarInit(&drained[threadIdx.x][o], 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint64_t stride = (uint64_t)gridDim.x * PAIRS;
    const uint64_t tile0 = (uint64_t)blockIdx.x * PAIRS + pair;
    const unsigned numMine = tile0 < numTiles ? (unsigned)((numTiles - tile0 + stride - 1) / stride) : 0u;

    if (mover) {
        // element (lane | i << 5), i < 16: lane part in registers, i part = map (constant bank); slab offsets combine by XOR
        uint64_t laneOff = 0;
        unsigned laneRow = 0, laneN = 0;
#pragma unroll
        for (unsigned b = 0; b < 5; b++) {
            const unsigned bit = (lane >> b) & 1u, role = localPos.pos[b];
            laneOff |= (uint64_t)bit << tileSpec.pos[b];
            if (role < T) laneRow |= bit << role; else laneN |= bit << (role - T);
        }
        char* laneG = reinterpret_cast<char*>(amps) + (laneOff << 4);
        const unsigned laneS = spec5Offset(laneRow, laneN, swz);
        auto load = [&](unsigned i) {                                  // tile i of this pair -> input slab i % NIN
            const char* src = laneG + (insertZeroBitsN<T + F>(tile0 + (uint64_t)i * stride, tileSpec) << 4);
            const unsigned s = i % NIN, dst = (in0 + s * SLAB_BYTES) ^ laneS;
            if (swz.viaL1) {
#pragma unroll
                for (unsigned e = 0; e < 16; e++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst ^ map.sByte[e]), "l"(src + map.gByte[e]) : "memory");
            } else {
#pragma unroll
                for (unsigned e = 0; e < 16; e++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst ^ map.sByte[e]), "l"(src + map.gByte[e]) : "memory");
            }
            cpAsyncArriveNoinc(&full[pair][s]);
        };
        for (unsigned i = 0; i < NIN && i < numMine; i++) load(i);
        for (unsigned i = 0; i < numMine; i++) {
            const unsigned o = i % NOUT;
            mbarWait(&done[pair][o], (i / NOUT) & 1u);
            char* dst = laneG + (insertZeroBitsN<T + F>(tile0 + (uint64_t)i * stride, tileSpec) << 4);
            const unsigned src = (out0 + o * SLAB_BYTES) ^ laneS;
#pragma unroll
            for (unsigned e0 = 0; e0 < 16; e0 += 8) {
                double2 v[8];
#pragma unroll
                for (unsigned e = 0; e < 8; e++)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[e].x), "=d"(v[e].y) : "r"(src ^ map.sByte[e0 + e]) : "memory");
#pragma unroll
                for (unsigned e = 0; e < 8; e++) *reinterpret_cast<double2*>(dst + map.gByte[e0 + e]) = v[e];
            }
            __syncwarp();                                               // every lane holds its share of the slab in registers
            if (lane == 0) mbarArrive(&drained[pair][o]);
            if (i + NIN < numMine) load(i + NIN);                       // input slab i % NIN was released by done
        }
        return;
    }

    // ---- compute warps: A-fragments of this warp's 16 gate rows (3M form), then LDS -> DMMA -> STS per tile
    const unsigned g = lane >> 2, q = lane & 3u;
    double ar[2][8], ad[2][8], as[2][8];
#pragma unroll
    for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const double2 e = gate[(16 * h + g + 8 * (v & 1)) * D + 16 * kb + q + 4 * (v >> 1)];
            ar[kb][v] = e.x;
            ad[kb][v] = e.y - e.x;
            as[kb][v] = e.x + e.y;
        }
    // thread parts of the operand (row q, column g) and result (row 16h + g, column 2q) offsets
    const unsigned laneB = spec5Offset(q, g, swz), laneC = spec5Offset(16 * h + g, 2 * q, swz);
    for (unsigned i = 0; i < numMine; i++) {
        const unsigned s = i % NIN, o = i % NOUT;
        const unsigned xB = (in0 + s * SLAB_BYTES) ^ laneB, yC = (out0 + o * SLAB_BYTES) ^ laneC;
        mbarWait(&full[pair][s], (i / NIN) & 1u);
        if (i >= NOUT) mbarWait(&drained[pair][o], ((i / NOUT) - 1u) & 1u);
#pragma unroll(UNROLL_NB ? 2 : 1)
        for (unsigned nb = 0; nb < VEC / 8; nb++) {
            double k1[4] = {0.0, 0.0, 0.0, 0.0}, k2[4] = {0.0, 0.0, 0.0, 0.0}, k3[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (unsigned kb = 0; kb < 2; kb++) {
                double xr[4], xi[4], xs[4];
#pragma unroll
                for (unsigned v = 0; v < 4; v++) {                      // b[v]: k = q + 4v (row 16kb + q + 4v), n = nb*8 + g
                    const unsigned u = spec5Offset(16 * kb + 4 * v, nb * 8, swz);       // warp-uniform instruction part
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xr[v]), "=d"(xi[v]) : "r"(xB ^ u) : "memory");
                    xs[v] = xr[v] + xi[v];
                }
                // the product that needs the DADD results goes last: the adds retire behind the other two products' DMMAs
                dmma16816(k2, ad[kb], xr);
                dmma16816(k3, as[kb], xi);
                dmma16816(k1, ar[kb], xs);
            }
#pragma unroll
            for (unsigned v = 0; v < 4; v++) {                          // c[v]: row 16h + g + 8(v>>1), column nb*8 + 2q + (v&1)
                const unsigned u = spec5Offset(8 * (v >> 1), nb * 8 + (v & 1), swz);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(yC ^ u), "d"(k1[v] - k3[v]), "d"(k1[v] + k2[v]) : "memory");
            }
        }
        __syncwarp();                                                   // all of this warp's reads and writes of the slabs are done
        if (lane == 0) mbarArrive(&done[pair][o]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// t == 6: 64 KiB of gate does not fit constant memory; block-wide tile with the gate transposed in shared memory,
// thread = (8 consecutive rows, one lane).
template <int R>
__global__ void __launch_bounds__(256) manyTargTileKernel(double2* amps, uint64_t numTiles, BitSpec tileSpec,
                                                         unsigned t, unsigned f, const double2* __restrict__ gateT, BitSpec localPos) {
    extern __shared__ double2 smem[];
    const unsigned d = 1u << t, lanes = 1u << f, tileAmps = d << f;
    double2* G = smem;                 // G^T[l][r], d*d
    double2* X = smem + (size_t)d * d; // X[row][lane]
    for (unsigned e = threadIdx.x; e < d * d; e += blockDim.x) G[e] = gateT[e];

    uint64_t gOff[8];
    unsigned xIdx[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned e = threadIdx.x + k * 256;
        uint64_t g = 0;
        unsigned row = 0, lane = 0;
        for (unsigned b = 0; b < t + f; b++) {
            unsigned bit = (e >> b) & 1u, role = localPos.pos[b];
            g |= (uint64_t)bit << tileSpec.pos[b];
            if (role < t) row |= bit << role; else lane |= bit << (role - t);
        }
        gOff[k] = g;
        xIdx[k] = row * lanes + lane;
    }
    const unsigned myLane = threadIdx.x % lanes, r0 = (threadIdx.x / lanes) * R;
    const bool active = r0 < d;

    for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const uint64_t base = insertZeroBits(tile, tileSpec);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (threadIdx.x + k * 256 < tileAmps) X[xIdx[k]] = amps[base | gOff[k]];
        __syncthreads();
        double2 acc[R];
#pragma unroll
        for (int i = 0; i < R; i++) acc[i] = make_double2(0.0, 0.0);
        if (active) {
            for (unsigned l = 0; l < d; l++) {
                const double2 x = X[l * lanes + myLane];
                const double2* grow = G + (size_t)l * d + r0;
#pragma unroll
                for (int i = 0; i < R; i++) acc[i] = cfma(grow[i], x, acc[i]);
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int i = 0; i < R; i++) X[(r0 + i) * lanes + myLane] = acc[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (threadIdx.x + k * 256 < tileAmps) amps[base | gOff[k]] = X[xIdx[k]];
    }
}

// ---------------------------------------------------------------------------------------------------------
// t >= 7 (krausMap superoperators): one block per 2^t-amplitude group staged in shared memory,
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

// i part of element (lane | i << 5): tile bits 5.. of the element index
template <int T>
void fillTileMap(const BitSpec& tileSpec, const BitSpec& localPos, TileMap<T>* map) {
    constexpr unsigned F = DmmaGeom<T>::F, S = DmmaGeom<T>::S;
    for (unsigned i = 0; i < DmmaGeom<T>::EPL; i++) {
        uint64_t off = 0;
        unsigned row = 0, n = 0;
        for (unsigned b = 5; b < T + F; b++) {
            const unsigned bit = (i >> (b - 5)) & 1u, role = localPos.pos[b];
            off |= (uint64_t)bit << tileSpec.pos[b];
            if (role < (unsigned)T) row |= bit << role; else n |= bit << (role - T);
        }
        map->gByte[i] = off << 4;
        map->sByte[i] = (row * S + n) << 4;
    }
}

template <int T, bool BULK>
int launchDmmaKernelImpl(dfsa_state* s, uint64_t numTiles, const BitSpec& tileSpec, const BitSpec& localPos, const double2* devGate, const TileMap<T>& map) {
    DfsaContext& ctx = dfsaCtx();
    constexpr unsigned WARPS = DmmaGeom<T>::WARPS;
    const size_t smemBytes = DmmaGeom<T>::smemBytes;                // 144 KiB (t=5), 102 KiB (t=4), 51 KiB (t=3)
    static int blocksPerSM = 0;
    if (blocksPerSM == 0) {
        DFSA_CUDA(cudaFuncSetAttribute(manyTargDmmaKernel<T, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, manyTargDmmaKernel<T, BULK>, 32 * WARPS, smemBytes) != cudaSuccess || blocksPerSM < 1)
            blocksPerSM = 1;
    }
    const uint64_t blocksNeeded = (numTiles + WARPS - 1) / WARPS;
    const unsigned grid = (unsigned)std::min<uint64_t>(blocksNeeded, (uint64_t)ctx.numSMs * blocksPerSM);
    manyTargDmmaKernel<T, BULK><<<grid, 32 * WARPS, smemBytes, ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, localPos, devGate, map);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

// full tiles only (f == DmmaGeom<T>::F free bits): shards too small for one are the caller's business
template <int T>
int launchDmmaKernel(dfsa_state* s, const uint32_t* targets, uint64_t targMask, const double2* devGate) {
    constexpr unsigned F = DmmaGeom<T>::F;
    const unsigned L = s->logNumAmps;
    BitSpec tileSpec, localPos;
    DFSA_TRY(buildTile(targets, T, L, targMask, F, &tileSpec, &localPos));
    TileMap<T> map;
    fillTileMap<T>(tileSpec, localPos, &map);
    // bulk rows need the VEC vectors of a tile row to be one contiguous run: free bits = address bits 0..F-1
    bool contiguous = !getenv("DFSA_MANYTARG_NO_BULK");
    for (unsigned b = 0; b < F && contiguous; b++) contiguous = (tileSpec.pos[b] == b) && (localPos.pos[b] >= (unsigned)T);
    const uint64_t numTiles = s->numAmps >> (T + F);
    return contiguous ? launchDmmaKernelImpl<T, true>(s, numTiles, tileSpec, localPos, devGate, map)
                      : launchDmmaKernelImpl<T, false>(s, numTiles, tileSpec, localPos, devGate, map);
}

// t == 5, warp-pair 3M kernel (every target placement; full tiles only)
int launchPair5(dfsa_state* s, const uint32_t* targets, uint64_t targMask, const double2* devGate) {
    DfsaContext& ctx = dfsaCtx();
    const unsigned L = s->logNumAmps;
    BitSpec tileSpec, localPos;
    DFSA_TRY(buildTile(targets, Pair5::T, L, targMask, Pair5::F, &tileSpec, &localPos));
    Pair5Map map;                                                    // j part of element (lane | h << 5 | j << 6): tile bits 6..8
    for (unsigned j = 0; j < Pair5::EPW; j++) {
        uint64_t off = 0;
        unsigned row = 0, n = 0;
        for (unsigned b = 6; b < Pair5::T + Pair5::F; b++) {
            const unsigned bit = (j >> (b - 6)) & 1u, role = localPos.pos[b];
            off |= (uint64_t)bit << tileSpec.pos[b];
            if (role < Pair5::T) row |= bit << role; else n |= bit << (role - Pair5::T);
        }
        map.gByte[j] = off << 4;
        map.sByte[j] = (row * Pair5::S + n) << 4;
    }
    static bool configured = false;
    if (!configured) {
        DFSA_CUDA(cudaFuncSetAttribute(manyTarg5PairKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pair5::smemBytes));
        configured = true;
    }
    const uint64_t numTiles = s->numAmps >> (Pair5::T + Pair5::F);
    const uint64_t blocksNeeded = (numTiles + Pair5::PAIRS - 1) / Pair5::PAIRS;
    const unsigned grid = (unsigned)std::min<uint64_t>(blocksNeeded, (uint64_t)ctx.numSMs);
    manyTarg5PairKernel<<<grid, 64 * Pair5::PAIRS, Pair5::smemBytes, ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, localPos, devGate, map);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

// Swizzle constants of the warp-specialised kernel: c[0], c[1] are fixed by the fragment access patterns; c[2..4] are
// picked so that the column fields of the eight lanes of a quarter-warp of the mover (address bits = tile bits 0..2,
// each either column bit n_j -> value 1 << j, or gate-row bit i -> value c[i]) are linearly independent over GF(2),
// i.e. the eight 16-byte accesses of a phase fall into eight different bank groups.
void chooseSpec5Swizzle(const BitSpec& localPos, Spec5Swz* z) {
    z->c[0] = 5; z->c[1] = 6; z->c[2] = z->c[3] = z->c[4] = 0;
    bool inSpan[8] = {true, false, false, false, false, false, false, false};
    auto add = [&](unsigned v) { bool next[8]; for (unsigned x = 0; x < 8; x++) next[x] = inSpan[x] || inSpan[x ^ v]; for (unsigned x = 0; x < 8; x++) inSpan[x] = next[x]; };
    unsigned pending[3], numPending = 0;
    for (unsigned b = 0; b < 3; b++) {
        const unsigned role = localPos.pos[b];
        if (role >= Spec5::T) { if (role - Spec5::T < 3) add(1u << (role - Spec5::T)); }   // column bit 3 does not select a bank group
        else if (role < 2) add(z->c[role]);
        else pending[numPending++] = role;
    }
    for (unsigned k = 0; k < numPending; k++)
        for (unsigned v = 1; v < 8; v++)
            if (!inSpan[v]) { z->c[pending[k]] = v; add(v); break; }
    for (unsigned i = 2, v = 1; i < 5; i++)                          // rows outside the quarter-warp bits: any value
        if (z->c[i] == 0) { z->c[i] = v; v = (v % 7) + 1; }
}

// t == 5, warp-specialised kernel (mover warps + compute warp pairs); same tile geometry as the one-warp kernel
template <unsigned NIN, bool UNROLL_NB>
int launchSpec5Impl(dfsa_state* s, uint64_t numTiles, const BitSpec& tileSpec, const BitSpec& localPos, const double2* devGate, const TileMap<5>& map, const Spec5Swz& swz) {
    DfsaContext& ctx = dfsaCtx();
    static bool configured = false;
    if (!configured) {
        DFSA_CUDA(cudaFuncSetAttribute(manyTarg5SpecKernel<NIN, UNROLL_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Spec5::smemBytes(NIN)));
        configured = true;
    }
    const uint64_t blocksNeeded = (numTiles + Spec5::PAIRS - 1) / Spec5::PAIRS;
    const unsigned grid = (unsigned)std::min<uint64_t>(blocksNeeded, (uint64_t)ctx.numSMs);
    manyTarg5SpecKernel<NIN, UNROLL_NB><<<grid, Spec5::THREADS, Spec5::smemBytes(NIN), ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, localPos, devGate, map, swz);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

int launchSpec5(dfsa_state* s, const uint32_t* targets, uint64_t targMask, const double2* devGate) {
    const unsigned L = s->logNumAmps;
    BitSpec tileSpec, localPos;
    DFSA_TRY(buildTile(targets, Spec5::T, L, targMask, Spec5::F, &tileSpec, &localPos));
    Spec5Swz swz;
    chooseSpec5Swizzle(localPos, &swz);
    // When address bit 0 is a gate-row bit, the two amplitudes of a 32-byte sector go to different slab rows and the
    // L1-bypassing cp.async.cg fetches the sector from L2 once per half (ncu: 2.5x the L2 read sectors, low targets);
    // routed through L1 (.ca) the second half hits. Data is read once per kernel, so L1 residency costs nothing else.
    swz.viaL1 = (localPos.pos[0] < Spec5::T) ? 1u : 0u;
    if (const char* e = getenv("DFSA_SPEC5_CA")) swz.viaL1 = (unsigned)atoi(e);
    // i part of element (lane | i << 5): shard byte offset and slab offset (row and swizzled column fields)
    TileMap<5> map;
    for (unsigned i = 0; i < 16; i++) {
        uint64_t off = 0;
        unsigned row = 0, n = 0;
        for (unsigned b = 5; b < Spec5::T + Spec5::F; b++) {
            const unsigned bit = (i >> (b - 5)) & 1u, role = localPos.pos[b];
            off |= (uint64_t)bit << tileSpec.pos[b];
            if (role < Spec5::T) row |= bit << role; else n |= bit << (role - Spec5::T);
        }
        map.gByte[i] = off << 4;
        map.sByte[i] = spec5Offset(row, n, swz);
    }
    const uint64_t numTiles = s->numAmps >> (Spec5::T + Spec5::F);
    const char* e = getenv("DFSA_SPEC5_NIN");
    const char* u = getenv("DFSA_SPEC5_UNROLL");
    const bool nin3 = e && atoi(e) == 3, unroll = !(u && atoi(u) == 0);
    if (nin3) return unroll ? launchSpec5Impl<3, true>(s, numTiles, tileSpec, localPos, devGate, map, swz)
                            : launchSpec5Impl<3, false>(s, numTiles, tileSpec, localPos, devGate, map, swz);
    return unroll ? launchSpec5Impl<4, true>(s, numTiles, tileSpec, localPos, devGate, map, swz)
                  : launchSpec5Impl<4, false>(s, numTiles, tileSpec, localPos, devGate, map, swz);
}

// DFSA_MANYTARG5=warp selects the one-warp-per-tile 4M kernel for t = 5 (kept for comparison runs); default is the pair kernel
int variant5() {                                                     // 0 = pair (default), 1 = warp, 2 = spec
    const char* e = getenv("DFSA_MANYTARG5");
    if (e && strcmp(e, "warp") == 0) return 1;
    if (e && strcmp(e, "spec") == 0) return 2;
    return 0;
}

}  // namespace

extern "C" int dfsa_k_manyTarg(dfsa_state* s, const uint32_t* targets, unsigned numTargets, const double* gate) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && targets && gate, "null argument");
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

    // tensor-core tiles need t + F local bits; a smaller shard (< 2^10 amplitudes) goes to the generic kernel below
    if (t >= 3 && t <= 5 && L >= t + (t == 5 ? DmmaGeom<5>::F : DmmaGeom<3>::F)) {
        void* stage; int slot;
        DFSA_TRY(dfsaStagingAcquire(gateBytes, &stage, &slot));
        memcpy(stage, gate, gateBytes);
        double2* dev;
        DFSA_TRY(dfsaScratch(gateBytes, &dev));
        DFSA_CUDA(cudaMemcpyAsync(dev, stage, gateBytes, cudaMemcpyHostToDevice, ctx.compute));
        DFSA_TRY(dfsaStagingCommit(slot));
        switch (t) {
            case 3:  return launchDmmaKernel<3>(s, targets, targMask, dev);
            case 4:  return launchDmmaKernel<4>(s, targets, targMask, dev);
            default:
                switch (variant5()) {
                    case 1:  return launchDmmaKernel<5>(s, targets, targMask, dev);
                    case 2:  return launchSpec5(s, targets, targMask, dev);
                    default: return launchPair5(s, targets, targMask, dev);
                }
        }
    }

    if (t == 6) {
        void* stage; int slot;
        DFSA_TRY(dfsaStagingAcquire(gateBytes, &stage, &slot));
        double2* gt = (double2*)stage;                    // transposed: G^T[l][r]
        for (uint64_t r = 0; r < d; r++) for (uint64_t l = 0; l < d; l++) gt[l * d + r] = hostAmp(gate + 2 * (r * d + l));
        double2* dev;
        DFSA_TRY(dfsaScratch(gateBytes, &dev));
        DFSA_CUDA(cudaMemcpyAsync(dev, stage, gateBytes, cudaMemcpyHostToDevice, ctx.compute));
        DFSA_TRY(dfsaStagingCommit(slot));
        const unsigned f = std::min(5u, L - t);
        BitSpec tileSpec, localPos;
        DFSA_TRY(buildTile(targets, t, L, targMask, f, &tileSpec, &localPos));
        const size_t smemBytes = (d * d + (d << f)) * sizeof(double2);
        const uint64_t numTiles = s->numAmps >> (t + f);
        const unsigned grid = (unsigned)std::min<uint64_t>(numTiles, (uint64_t)ctx.numSMs * 2);
        DFSA_CUDA(cudaFuncSetAttribute(manyTargTileKernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        manyTargTileKernel<8><<<grid, 256, smemBytes, ctx.compute>>>(s->arr[DFSA_AMPS], numTiles, tileSpec, t, f, dev, localPos);
        DFSA_LAUNCH_CHECK();
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
