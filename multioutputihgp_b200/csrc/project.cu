// K-project: OILMM projection of the observations onto the L orthogonal latents (sm_100a).
//
// Replaces (reference, /root/reference/moihgp/include/moihgp/moihgp.h):
//   :159-182  Ty = diag(S^-1/2) U' y         (and the identical blocks at :240-263, :314-336, :471-498, :625-648)
//   :499-501  || (I - U U') y ||_2           (norm, not squared: Q9)
//   :222-225  yhat = U diag(sqrt S) Tyhat    (back-projection, k_backproject)
//
// Layouts.  Y is the caller's [N][T][p] (time-major, like the reference's vector<VectorXd>).
// The projected series is written LATENT-MAJOR, u[n][l][t], so that the scan kernels read each
// latent's time series with unit stride.  One thread owns one time step (a row of Y) and LB
// latents at a time; the Y tile and the U panel are staged through shared memory with coalesced
// loads (row pitch padded by one double: conflict-free row-per-lane reads).
#include <cuda_runtime.h>
#include <math.h>
#include "moihgp_device.cuh"
#include "launch.h"

namespace moihgp {

namespace {

constexpr int PT = 128;   // time steps (threads) per CTA
constexpr int PC = 16;    // columns of Y per staged panel
constexpr int LB = 8;     // latents accumulated in registers per pass

// grid: ceil(T / PT) * N CTAs (tile-minor).  Outputs:
//   u[n][l][t]    = S_l^-1/2 * sum_r U[r][l] y[r]          (always)
//   w[n][l][t]    = sum_r U[r][l] y[r]                      (optional, objective path)
//   yl[n][l][t]   = y[l]  (raw output l, for pv: moihgp.h:510, Q8)   (optional, objective path)
//   rho[n][t]     = || y - U U' y ||_2                      (optional)
//   nan_flag      = set to 1 if any y is NaN (missing-data rows need the LS projection path)
__global__ void __launch_bounds__(PT) k_project(const double* __restrict__ Y, const double* __restrict__ U,
                                               const double* __restrict__ S, int p, int L, long long T,
                                               double* __restrict__ u, double* __restrict__ w, double* __restrict__ yl,
                                               double* __restrict__ rho, int* __restrict__ nan_flag) {
    extern __shared__ double sm[];
    double* ys = sm;                       // [PT][PC + 1]
    double* us = ys + PT * (PC + 1);       // [PC][L]
    double* ws = us + PC * L;              // [PT][L + 1]   projected (unscaled) values of this tile
    const int tid = threadIdx.x;
    const long long tiles = (T + PT - 1) / PT;
    const long long n = blockIdx.x / tiles;
    const long long t0 = (blockIdx.x - n * tiles) * PT;
    const long long t = t0 + tid;
    const bool live = t < T;
    const double* Yn = Y + n * T * p;
    const int rows = (int)min((long long)PT, T - t0);
    bool saw_nan = false;

    // ---- pass 1: w = U' y, LB latents at a time -----------------------------------------------
    for (int l0 = 0; l0 < L; l0 += LB) {
        double acc[LB];
#pragma unroll
        for (int j = 0; j < LB; ++j) acc[j] = 0.0;
        for (int r0 = 0; r0 < p; r0 += PC) {
            const int pc = min(PC, p - r0);
            __syncthreads();
            for (int i = tid; i < rows * pc; i += PT) {
                const int row = i / pc, col = i - row * pc;
                ys[row * (PC + 1) + col] = Yn[(t0 + row) * p + r0 + col];
            }
            for (int i = tid; i < pc * L; i += PT) us[i] = U[(size_t)r0 * L + i];
            __syncthreads();
            if (live) {
                for (int c = 0; c < pc; ++c) {
                    const double y = ys[tid * (PC + 1) + c];
                    if (l0 == 0) saw_nan = saw_nan || isnan(y);
#pragma unroll
                    for (int j = 0; j < LB; ++j)
                        if (l0 + j < L) acc[j] = fma(us[c * L + l0 + j], y, acc[j]);
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < LB; ++j)
                if (l0 + j < L) {
                    const int l = l0 + j;
                    ws[tid * (L + 1) + l] = acc[j];
                    const size_t o = ((size_t)n * L + l) * T + t;
                    u[o] = acc[j] * (1.0 / sqrt(S[l]));
                    if (w) w[o] = acc[j];
                }
        }
    }
    if (saw_nan) *nan_flag = 1;

    // ---- pass 2: residual norm and raw y(l) ---------------------------------------------------
    if (rho || yl) {
        double q = 0.0;
        for (int r0 = 0; r0 < p; r0 += PC) {
            const int pc = min(PC, p - r0);
            __syncthreads();
            for (int i = tid; i < rows * pc; i += PT) {
                const int row = i / pc, col = i - row * pc;
                ys[row * (PC + 1) + col] = Yn[(t0 + row) * p + r0 + col];
            }
            for (int i = tid; i < pc * L; i += PT) us[i] = U[(size_t)r0 * L + i];
            __syncthreads();
            if (live) {
                for (int c = 0; c < pc; ++c) {
                    const double y = ys[tid * (PC + 1) + c];
                    double e = y;
                    for (int l = 0; l < L; ++l) e = fma(-us[c * L + l], ws[tid * (L + 1) + l], e);
                    q = fma(e, e, q);
                    if (yl && r0 + c < L) yl[((size_t)n * L + (r0 + c)) * T + t] = y;
                }
            }
        }
        if (live && rho) rho[(size_t)n * T + t] = sqrt(q);
    }
}

// Back-projection of the filtered function values: Yhat[n][t][r] = sum_l U[r][l] sqrt(S_l) X[n][t][l][0]
// (ihgp.h:51 yhat = xnew(0); moihgp.h:222-225).  grid: (ceil(T / PT), N)
__global__ void __launch_bounds__(PT) k_backproject(const double* __restrict__ X, const double* __restrict__ U,
                                                   const double* __restrict__ S, int p, int L, int d, long long T,
                                                   double* __restrict__ Yhat) {
    extern __shared__ double sm[];
    double* us = sm;                 // [p][L] scaled by sqrt(S)
    double* fs = us + (size_t)p * L; // [PT][L + 1]
    const int tid = threadIdx.x;
    const long long tiles = (T + PT - 1) / PT;
    const long long n = blockIdx.x / tiles;
    const long long t0 = (blockIdx.x - n * tiles) * PT;
    const int rows = (int)min((long long)PT, T - t0);
    for (int i = tid; i < p * L; i += PT) us[i] = U[i] * sqrt(S[i % L]);
    const double* Xn = X + ((size_t)n * T + t0) * L * d;
    for (int i = tid; i < rows * L; i += PT) fs[(i / L) * (L + 1) + (i % L)] = Xn[(size_t)i * d];
    __syncthreads();
    double* Yo = Yhat + ((size_t)n * T + t0) * p;
    for (int i = tid; i < rows * p; i += PT) {
        const int row = i / p, r = i - row * p;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s = fma(us[r * L + l], fs[row * (L + 1) + l], s);
        Yo[i] = s;
    }
}

}  // namespace

cudaError_t launch_project(const double* Y, const double* U, const double* S, int p, int L, long long N, long long T,
                           double* u, double* w, double* yl, double* rho, int* nan_flag, cudaStream_t stream) {
    const size_t smem = sizeof(double) * ((size_t)PT * (PC + 1) + (size_t)PC * L + (size_t)PT * (L + 1));
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)(((T + PT - 1) / PT) * N);
    k_project<<<grid, PT, smem, stream>>>(Y, U, S, p, L, T, u, w, yl, rho, nan_flag);
    return cudaGetLastError();
}

cudaError_t launch_backproject(const double* X, const double* U, const double* S, int p, int L, int d, long long N,
                               long long T, double* Yhat, cudaStream_t stream) {
    const size_t smem = sizeof(double) * ((size_t)p * L + (size_t)PT * (L + 1));
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_backproject, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)(((T + PT - 1) / PT) * N);
    k_backproject<<<grid, PT, smem, stream>>>(X, U, S, p, L, d, T, Yhat);
    return cudaGetLastError();
}

}  // namespace moihgp
