// Microbenchmark: the inner loop of k_project_rows in isolation (operands from shared memory, no global traffic).
// What DMMA rate does "2 LDS (Y) + NB LDS (U) + 2 DFMA + 2 NB DMMA per k block" sustain at W warps per SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop dmma_loop.cu ; run on the B200 box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE bit 0: DFMA of y^2 inline; bit 1: conflict-free Y rows (row = 2 g4 + rb instead of g4 + 8 rb); bit 2: no U loads (register B)
template <int NB, int RBN, int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, int kblocks) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5, g4 = lane >> 2, q4 = lane & 3;
    constexpr int UP = 8 * NB + 4;
    double* usm = sm;                       // [64][UP]
    double* ysm = sm + 64 * UP + wi * 1024; // per warp: 4 boxes of [16][16]
    for (int i = tid; i < 64 * UP; i += blockDim.x) usm[i] = 1e-3 * (i % 7);
    for (int i = lane; i < 1024; i += 32) ysm[i] = 1e-3 * (i % 5);
    __syncthreads();
    unsigned yoff[RBN][4];
    for (int rb = 0; rb < RBN; ++rb)
        for (int k3 = 0; k3 < 4; ++k3) {
            const int row = (MODE & 2) ? (2 * g4 + rb) & 15 : (g4 + 8 * rb) & 15;
            yoff[rb][k3] = row * 128 + ((((2 * k3 + (q4 >> 1)) ^ row) & 7) << 4) + ((q4 & 1) << 3);
        }
    double acc[RBN][NB][2];
    for (int rb = 0; rb < RBN; ++rb) for (int nb = 0; nb < NB; ++nb) { acc[rb][nb][0] = 0; acc[rb][nb][1] = 0; }
    double sy[RBN] = {};
    double ureg[NB];
    for (int nb = 0; nb < NB; ++nb) ureg[nb] = 1e-3 * nb;
    const unsigned char* ys = reinterpret_cast<const unsigned char*>(ysm);
    for (int it = 0; it < iters; ++it) {
        const double* ub = usm + q4 * UP + g4;
        for (int j = 0; j < kblocks / 4; ++j) {
#pragma unroll
            for (int k3 = 0; k3 < 4; ++k3) {
                double yv[RBN], uv[NB];
#pragma unroll
                for (int rb = 0; rb < RBN; ++rb) {
                    yv[rb] = *reinterpret_cast<const double*>(ys + (j & 3) * 2048 + yoff[rb][k3]);
                    if (MODE & 1) sy[rb] = fma(yv[rb], yv[rb], sy[rb]);
                }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) uv[nb] = (MODE & 4) ? ureg[nb] : ub[(size_t)((16 * j + 4 * k3) & 63) * UP + 8 * nb];
#pragma unroll
                for (int rb = 0; rb < RBN; ++rb)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) dmma(acc[rb][nb][0], acc[rb][nb][1], uv[nb], yv[rb]);
            }
        }
        __syncwarp();
    }
    double s = 0;
    for (int rb = 0; rb < RBN; ++rb) { s += sy[rb]; for (int nb = 0; nb < NB; ++nb) s += acc[rb][nb][0] + acc[rb][nb][1]; }
    out[blockIdx.x * blockDim.x + tid] = s;
}

template <int NB, int RBN, int MODE>
void run(int warps, const char* name) {
    const int iters = 2000, kblocks = 16, blocks = 148, threads = 32 * warps;
    const size_t smem = sizeof(double) * (64 * (8 * NB + 4) + (size_t)warps * 1024);
    cudaFuncSetAttribute(k<NB, RBN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NB, RBN, MODE><<<blocks, threads, smem>>>(out, 10, kblocks);
    cudaEventRecord(e0);
    k<NB, RBN, MODE><<<blocks, threads, smem>>>(out, iters, kblocks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)iters * kblocks * RBN * NB * blocks * warps;
    printf("%-46s NB=%d RB=%d warps/SM=%2d  %.3f ms  %.3e DMMA/s = %.1f%% of 7.23e10  err=%s\n", name, NB, RBN, warps, ms, n / ms * 1e3,
           100.0 * n / ms * 1e3 / 7.23e10, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int w : {8, 16}) {
        run<4, 2, 1>(w, "as in k_project_rows (DFMA inline, 4-way Y)");
        run<4, 2, 0>(w, "no DFMA");
        run<4, 2, 3>(w, "DFMA inline, conflict-free Y rows");
        run<4, 2, 2>(w, "no DFMA, conflict-free Y rows");
        run<4, 2, 6>(w, "no DFMA, conflict-free Y, U in registers");
        run<4, 4, 3>(w, "4 row blocks (32 rows/warp), DFMA, cf Y");
        run<8, 2, 3>(w, "NB=8 (L=64), DFMA, cf Y");
        run<2, 4, 3>(w, "NB=2, 4 row blocks, DFMA, cf Y");
    }
    return 0;
}
