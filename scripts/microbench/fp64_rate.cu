// Microbenchmark: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) issue rate on sm_100a, alone and mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu ; run on the B200 box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs): 2048 flop per warp instruction
__device__ __forceinline__ void dmma16(double (&c)[4], double a, double b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
}
__global__ void k16(double* out, int iters, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9 + i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma16(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void run16(int warps_per_sm) {
    const int iters = 20000, blocks = 148, threads = 32 * warps_per_sm;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k16<<<blocks, threads>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    k16<<<blocks, threads>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = 4.0 * iters * blocks * warps_per_sm;
    printf("DMMA16x8x8 warps/SM=%2d  %.3f ms  %.2f TFLOP/s  err=%s\n", warps_per_sm, ms, n * 2048 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

template <int MODE>   // 0: DFMA only, 1: DMMA only, 2: mixed (8 DFMA per DMMA)
__global__ void k(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = i * 0.5;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    const int iters = 20000;
    const int blocks = 148, threads = 32 * warps_per_sm;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * warps_per_sm;
    const double dfma = (MODE == 1 ? 0 : 64.0) * iters * warps;      // warp-level DFMA instructions
    const double dm = (MODE == 0 ? 0 : 8.0) * iters * warps;         // warp-level DMMA instructions
    const double flops = dfma * 32 * 2 + dm * 512;
    printf("%-6s warps/SM=%2d  %.3f ms  %.2f TFLOP/s  (DFMA %.3g/s, DMMA %.3g/s warp-instr)  err=%s\n", name, warps_per_sm, ms,
           flops / ms / 1e9, dfma / ms * 1e3, dm / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) { run<0>("DFMA", w); run<1>("DMMA", w); run<2>("MIXED", w); run16(w); }
    return 0;
}
