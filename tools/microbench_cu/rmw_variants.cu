// In-place read-modify-write streaming variants over a 16 GiB double2 array: what is the practical HBM ceiling for the
// oneTargGate access pattern (pairs 2^t apart) and does unroll depth / 256-bit access / block size matter?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t ins0(uint64_t v, unsigned p) { uint64_t m = (1ULL << p) - 1; return ((v & ~m) << 1) | (v & m); }
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

struct G { double2 m00, m01, m10, m11; };

template <int TPB, int UNROLL>
__global__ void __launch_bounds__(TPB) pair128(double2* a, uint64_t items, unsigned t, G g) {
    const uint64_t chunk = (uint64_t)TPB * UNROLL, bit = 1ULL << t;
    for (uint64_t base = blockIdx.x * chunk; base < items; base += (uint64_t)gridDim.x * chunk) {
        double2 x0[UNROLL], x1[UNROLL]; uint64_t i0[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { uint64_t j = base + u * TPB + threadIdx.x; i0[u] = ins0(j, t); x0[u] = a[i0[u]]; x1[u] = a[i0[u] | bit]; }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { a[i0[u]] = cfma(g.m01, x1[u], cmul(g.m00, x0[u])); a[i0[u] | bit] = cfma(g.m11, x1[u], cmul(g.m10, x0[u])); }
    }
}

struct D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ld256(const double2* p) { D4 r; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p)); return r; }
__device__ __forceinline__ void st256(double2* p, D4 v) { asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory"); }

// each item = two ADJACENT pairs (bit 0 free): 32-byte accesses; requires t >= 1
template <int TPB, int UNROLL>
__global__ void __launch_bounds__(TPB) pair256(double2* a, uint64_t items /* A/4 */, unsigned t, G g) {
    const uint64_t chunk = (uint64_t)TPB * UNROLL, bit = 1ULL << t;
    for (uint64_t base = blockIdx.x * chunk; base < items; base += (uint64_t)gridDim.x * chunk) {
        D4 x0[UNROLL], x1[UNROLL]; uint64_t i0[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { uint64_t j = base + u * TPB + threadIdx.x; i0[u] = ins0(j << 1, t); x0[u] = ld256(a + i0[u]); x1[u] = ld256(a + (i0[u] | bit)); }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            double2 p0 = make_double2(x0[u].a, x0[u].b), p1 = make_double2(x1[u].a, x1[u].b), q0 = make_double2(x0[u].c, x0[u].d), q1 = make_double2(x1[u].c, x1[u].d);
            double2 r0 = cfma(g.m01, p1, cmul(g.m00, p0)), r1 = cfma(g.m11, p1, cmul(g.m10, p0)), s0 = cfma(g.m01, q1, cmul(g.m00, q0)), s1 = cfma(g.m11, q1, cmul(g.m10, q0));
            st256(a + i0[u], D4{r0.x, r0.y, s0.x, s0.y}); st256(a + (i0[u] | bit), D4{r1.x, r1.y, s1.x, s1.y});
        }
    }
}

template <class F> float timeIt(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < reps; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}

int main() {
    const unsigned nq = 30; const uint64_t A = 1ULL << nq;
    double2* a; cudaMalloc(&a, A * sizeof(double2)); cudaMemset(a, 0, A * sizeof(double2));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    G g{{0.6, 0.1}, {0.3, -0.2}, {-0.3, 0.2}, {0.6, -0.1}};
    for (unsigned t : {1u, 5u, 12u, 29u}) {
#define RUN128(TPB, U, BPS) { int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair128<TPB, U>, TPB, 0); int grid = sms * (BPS ? BPS : nb); \
        float ms = timeIt([&] { pair128<TPB, U><<<grid, TPB>>>(a, A / 2, t, g); }, 5); printf("{\"variant\":\"pair128 tpb=%d unroll=%d bps=%d\",\"t\":%u,\"GBps\":%.1f}\n", TPB, U, BPS ? BPS : nb, t, 32.0 * A / ms / 1e6); }
#define RUN256(TPB, U, BPS) { int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair256<TPB, U>, TPB, 0); int grid = sms * (BPS ? BPS : nb); \
        float ms = timeIt([&] { pair256<TPB, U><<<grid, TPB>>>(a, A / 4, t, g); }, 5); printf("{\"variant\":\"pair256 tpb=%d unroll=%d bps=%d\",\"t\":%u,\"GBps\":%.1f}\n", TPB, U, BPS ? BPS : nb, t, 32.0 * A / ms / 1e6); }
        RUN128(256, 2, 0) RUN128(256, 4, 0) RUN128(512, 2, 0) RUN128(256, 1, 0) RUN128(128, 2, 0) RUN128(256, 2, 4)
        RUN256(256, 1, 0) RUN256(256, 2, 0) RUN256(128, 2, 0) RUN256(512, 1, 0) RUN256(256, 2, 4) RUN256(256, 4, 0)
    }
    // out-of-place copy reference on the same box
    double2* b; cudaMalloc(&b, (A / 2) * sizeof(double2));
    float ms = timeIt([&] { cudaMemcpyAsync(b, a, (A / 2) * sizeof(double2), cudaMemcpyDeviceToDevice); }, 5);
    printf("{\"variant\":\"cudaMemcpy D2D 8 GiB\",\"GBps\":%.1f}\n", 2.0 * (A / 2) * 16 / ms / 1e6);
    return 0;
}
