// Measures the FP64 peaks of the device: plain DFMA (register operands) and DMMA m16n8k16 / m8n8k4 (mma.sync f64).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfmaKernel(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-9 + i;
    double x = 1.0000001, y = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816Kernel(double* out, int iters) {
    double a[8], b[4], c[4][4];
    for (int i = 0; i < 8; i++) a[i] = 1e-3 * (threadIdx.x + i);
    for (int i = 0; i < 4; i++) b[i] = 1e-3 * (threadIdx.x - i);
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) c[j][i] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0; for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma884Kernel(double* out, int iters) {
    double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x + 1), c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K> double timeIt(K launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int warpsPerSM : {4, 8, 16, 32}) {
        int blocks = sms * warpsPerSM / 4, threads = 128, iters = 20000;
        double ms = timeIt([&] { dfmaKernel<<<blocks, threads>>>(out, iters); });
        double flop = 2.0 * 16 * iters * (double)blocks * threads;
        printf("{\"kernel\":\"DFMA\",\"warps_per_sm\":%d,\"TFLOPs\":%.2f}\n", warpsPerSM, flop / ms / 1e9);
        ms = timeIt([&] { dmma16816Kernel<<<blocks, threads>>>(out, iters / 4); });
        flop = 2.0 * 16 * 8 * 16 * 4 * (iters / 4) * (double)blocks * (threads / 32);
        printf("{\"kernel\":\"DMMA.m16n8k16\",\"warps_per_sm\":%d,\"TFLOPs\":%.2f}\n", warpsPerSM, flop / ms / 1e9);
        ms = timeIt([&] { dmma884Kernel<<<blocks, threads>>>(out, iters); });
        flop = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
        printf("{\"kernel\":\"DMMA.m8n8k4\",\"warps_per_sm\":%d,\"TFLOPs\":%.2f}\n", warpsPerSM, flop / ms / 1e9);
    }
    return 0;
}
