// K-project: OILMM projection of the observations onto the L orthogonal latents (sm_100a).
//
// Replaces (reference, /root/reference/moihgp/include/moihgp/moihgp.h):
//   :159-182  Ty = diag(S^-1/2) U' y         (and the identical blocks at :240-263, :314-336, :471-498, :625-648)
//   :499-501  || (I - U U') y ||_2           (norm, not squared: Q9)
//   :222-225  yhat = U diag(sqrt S) Tyhat    (back-projection, k_backproject)
//
// Layouts.  Y is the caller's [N][T][p] (time-major, like the reference's vector<VectorXd>).
// The projected series is written LATENT-MAJOR, u[n][l][t], so that the scan kernels read each
// latent's time series with unit stride.  The residual norms are only ever summed over time
// (moihgp.h:503,563 through the callers' loops), so each CTA (a tile of PT = 128 time steps of one
// sequence) writes ONE partial sum rho_part[n][tile], reduced later in fixed order.
//
// k_project_mma<NB>: W[128 x L] = Ytile[128 x p] * U[p x L] on the FP64 tensor pipe (DMMA m8n8k4), K panels of 16
//   columns double-buffered through shared memory by cp.async; 8 warps x 16 rows, accumulators in registers
//   (L <= 8 NB).  The residual norm comes from ||y||^2 - ||U'y||^2 (U has orthonormal columns); rows where that
//   difference cancels (< 1e-4 ||y||^2) or is NaN are re-evaluated explicitly.  Needs p even (16-byte rows).
// k_project: the scalar-FMA version for every other shape.
#include <cuda_runtime.h>
#include <math.h>
#include "moihgp_device.cuh"
#include "launch.h"
#include "ls_project.cuh"

namespace moihgp {

namespace {

constexpr int PT = 128;   // time steps per CTA (both kernels)
constexpr int PC = 16;    // columns of Y per staged panel
constexpr int LB = 8;     // latents accumulated in registers per pass (scalar kernel)
constexpr unsigned FULL = 0xffffffffu;

// fixed-order sum of one value per thread over the CTA; result valid in thread 0
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < NT / 32; ++i) s += red[i];
    return s;
}

// grid: ceil(T / PT) * N CTAs (tile-minor).  Outputs:
//   u[n][l][t]    = S_l^-1/2 * sum_r U[r][l] y[r]          (always)
//   w[n][l][t]    = sum_r U[r][l] y[r]                      (optional, objective path)
//   yl[n][l][t]   = y[l]  (raw output l, for pv: moihgp.h:510, Q8)   (optional, objective path)
//   rho_part[n][tile] = sum over the tile of || y - U U' y ||_2      (optional)
//   nan_info      = {flag, count}: flag set to 1 if any y is NaN; rows with a NaN are appended to nan_rows (global row
//                   index n*T + t, at most nan_cap entries) and re-projected by k_nan_fix (missing-data LS projection)
__device__ __forceinline__ void note_nan_row(int* nan_info, long long* nan_rows, long long nan_cap, long long row) {
    nan_info[0] = 1;
    const int idx = atomicAdd(nan_info + 1, 1);
    if (idx < nan_cap) nan_rows[idx] = row;
    else nan_info[0] = 2;                    // overflow: reported by the host entry points
}

__global__ void __launch_bounds__(PT) k_project(const double* __restrict__ Y, const double* __restrict__ U,
                                               const double* __restrict__ S, int p, int L, long long T,
                                               double* __restrict__ u, double* __restrict__ w, double* __restrict__ yl,
                                               double* __restrict__ rho_part, int* __restrict__ nan_info,
                                               long long* __restrict__ nan_rows, long long nan_cap) {
    extern __shared__ double sm[];
    __shared__ double red[PT / 32];
    double* ys = sm;                       // [PT][PC + 1]
    double* us = ys + PT * (PC + 1);       // [PC][L]
    double* ws = us + PC * L;              // [PT][L + 1]   projected (unscaled) values of this tile
    const int tid = threadIdx.x;
    const long long tiles = (T + PT - 1) / PT;
    const long long n = blockIdx.x / tiles;
    const long long tile = blockIdx.x - n * tiles;
    const long long t0 = tile * PT;
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
    if (saw_nan) note_nan_row(nan_info, nan_rows, nan_cap, n * T + t);

    // ---- pass 2: residual norm and raw y(l) ---------------------------------------------------
    if (rho_part || yl) {
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
        if (rho_part) {
            const double s = block_sum<PT>(live ? sqrt(q) : 0.0, red);
            if (tid == 0) rho_part[n * tiles + tile] = s;
        }
    }
}

// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 16-byte cp.async with zero fill: `bytes` (0, 8 or 16) are read from gmem, the rest of the 16 bytes are zeros
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int MW = 8;                 // warps per CTA, 16 rows each
constexpr int MT = 32 * MW;           // threads
constexpr int KP = 16;                // columns of Y per K panel (4 k blocks)
constexpr int YROWB = KP * 8;         // bytes per row of a staged Y panel (128)

template <int NB>
struct MmaSmem {
    static constexpr int LP = 8 * NB;
    static constexpr int UPITCH = LP + 4;                   // doubles: the 4 k rows of a B fragment land in distinct bank groups
    static constexpr int YBYTES = PT * YROWB;               // one Y panel
    static constexpr int UBYTES = KP * UPITCH * 8;          // one U panel
    static constexpr int BYTES = 2 * YBYTES + 2 * UBYTES;
};

// grid: ceil(T / PT) * N CTAs (tile-minor), MT threads.  Same outputs as k_project.
template <int NB>
__global__ void __launch_bounds__(MT, NB <= 4 ? 3 : 2) k_project_mma(const double* __restrict__ Y, const double* __restrict__ U,
                                                   const double* __restrict__ S, int p, int L, long long T,
                                                   double* __restrict__ u, double* __restrict__ w, double* __restrict__ yl,
                                                   double* __restrict__ rho_part, int* __restrict__ nan_info,
                                                   long long* __restrict__ nan_rows, long long nan_cap) {
    using SMC = MmaSmem<NB>;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ double red[MW];
    __shared__ double rs_s[8 * NB];                 // S^-1/2 per latent (moihgp.h:181)
    __shared__ int any_bad;
    unsigned char* ysm = smraw;                                                   // [2][PT][KP] swizzled 16-byte chunks
    double* usm = reinterpret_cast<double*>(smraw + 2 * SMC::YBYTES);             // [2][KP][UPITCH]
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int g4 = lane >> 2, q4 = lane & 3;
    const long long tiles = (T + PT - 1) / PT;
    const long long n = blockIdx.x / tiles;
    const long long tile = blockIdx.x - n * tiles;
    const long long t0 = tile * PT;
    const int rows = (int)min((long long)PT, T - t0);
    const double* Yn = Y + ((size_t)n * T + t0) * p;
    const int npanels = (p + KP - 1) / KP;
    if (tid == 0) any_bad = 0;
    if (tid < 8 * NB) rs_s[tid] = tid < L ? 1.0 / sqrt(__ldg(S + tid)) : 0.0;
    const bool u_vec = L % 2 == 0 && (reinterpret_cast<size_t>(U) & 15) == 0;

    // stage panel kp: Y rows [0, PT) x columns [kp*KP, kp*KP + KP) and U rows [kp*KP, +KP) x [0, LP)
    auto stage = [&](int kp, int buf) {
        unsigned char* yb = ysm + buf * SMC::YBYTES;
        const int c0 = kp * KP;
#pragma unroll
        for (int i = 0; i < PT * (KP / 2) / MT; ++i) {
            const int q = tid + MT * i;
            const int row = q >> 3, ch = q & 7;                                   // 8 chunks of 2 columns per row
            const int col = c0 + 2 * ch;
            int bytes = row < rows ? max(0, min(2, p - col)) * 8 : 0;
            const double* src = Yn + (size_t)(row < rows ? row : 0) * p + (col < p ? col : 0);
            cp_async16_zfill(yb + row * YROWB + ((ch ^ (2 * (row & 3))) << 4), src, bytes);
        }
        double* ub = usm + buf * (SMC::UBYTES / 8);
        if (u_vec) {
            // U rows are 16-byte aligned (L even): asynchronous 16-byte copies, zero fill beyond p / L
            for (int i = tid; i < KP * (SMC::LP / 2); i += MT) {
                const int k = i / (SMC::LP / 2), l = 2 * (i - k * (SMC::LP / 2));
                const int bytes = (c0 + k < p && l < L) ? 16 : 0;
                cp_async16_zfill(ub + k * SMC::UPITCH + l, U + (size_t)(c0 + k < p ? c0 + k : 0) * L + (l < L ? l : 0), bytes);
            }
        } else {
            for (int i = tid; i < KP * SMC::LP; i += MT) {
                const int k = i / SMC::LP, l = i - k * SMC::LP;
                ub[k * SMC::UPITCH + l] = (c0 + k < p && l < L) ? __ldg(U + (size_t)(c0 + k) * L + l) : 0.0;
            }
        }
        cp_async_commit();
    };

    double acc[2][NB][2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { acc[rb][nb][0] = 0.0; acc[rb][nb][1] = 0.0; }
    double sy[2] = {0.0, 0.0};
    const int row0 = wi * 16 + g4;                                                // this lane's rows: row0, row0 + 8
    const long long tA = t0 + row0, tB = tA + 8;

    stage(0, 0);
    for (int kp = 0; kp < npanels; ++kp) {
        const int buf = kp & 1;
        if (kp + 1 < npanels) { stage(kp + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const unsigned char* yb = ysm + buf * SMC::YBYTES;
        const double* ub = usm + buf * (SMC::UBYTES / 8);
#pragma unroll
        for (int kb = 0; kb < KP / 4; ++kb) {
            double a[2], b[NB];
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
                const int row = row0 + 8 * rb;
                a[rb] = *reinterpret_cast<const double*>(yb + row * YROWB + (((2 * kb + (q4 >> 1)) ^ (2 * (row & 3))) << 4) + ((q4 & 1) << 3));
                sy[rb] = fma(a[rb], a[rb], sy[rb]);
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) b[nb] = ub[(4 * kb + q4) * SMC::UPITCH + 8 * nb + g4];
#pragma unroll
            for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) dmma884(acc[rb][nb][0], acc[rb][nb][1], a[rb], b[nb]);   // U' y   moihgp.h:181
            // raw y(l) for l < L (pv uses it: moihgp.h:510, Q8)
            if (yl) {
                const int col = kp * KP + 4 * kb + q4;
                if (col < L) {
                    if (tA < T) yl[((size_t)n * L + col) * T + tA] = a[0];
                    if (tB < T) yl[((size_t)n * L + col) * T + tB] = a[1];
                }
            }
        }
        __syncthreads();
    }

    // ---- outputs ---------------------------------------------------------------------------------
    double sw[2] = {0.0, 0.0};
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int l = 8 * nb + 2 * q4 + e;
            if (l < L) {
                const double rs = rs_s[l];
#pragma unroll
                for (int rb = 0; rb < 2; ++rb) {
                    const long long t = rb == 0 ? tA : tB;
                    const double c = acc[rb][nb][e];
                    sw[rb] = fma(c, c, sw[rb]);
                    if (t < T) {
                        const size_t o = ((size_t)n * L + l) * T + t;
                        u[o] = c * rs;
                        if (w) w[o] = c;
                    }
                }
            }
        }
    // residual norm of rows tA, tB: quad all-reduce of the two squared norms
    double rho_sum = 0.0;
    bool bad_row[2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        double ysq = sy[rb], wsq = sw[rb];
        ysq += __shfl_xor_sync(FULL, ysq, 1); ysq += __shfl_xor_sync(FULL, ysq, 2);
        wsq += __shfl_xor_sync(FULL, wsq, 1); wsq += __shfl_xor_sync(FULL, wsq, 2);
        const double q = ysq - wsq;
        const long long t = rb == 0 ? tA : tB;
        bad_row[rb] = t < T && !(q >= 1e-4 * ysq);
        if (q4 == 0 && t < T && !bad_row[rb]) rho_sum += sqrt(q);
    }
    if (bad_row[0] || bad_row[1]) any_bad = 1;
    __syncthreads();
    if (any_bad) {
        // explicit || y - U (U' y) ||_2 (moihgp.h:651) for the rows whose norm difference cancelled (or is NaN): one
        // lane of the quad per row, straight from global memory.  Rare.
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
            if (bad_row[rb] && q4 == 0) {
                const double* yr = Yn + (size_t)(row0 + 8 * rb) * p;
                double wl[8 * NB];
                bool isn = false;
                for (int l = 0; l < L; ++l) {
                    double a = 0.0;
                    for (int r = 0; r < p; ++r) a = fma(__ldg(U + (size_t)r * L + l), yr[r], a);
                    wl[l] = a;
                }
                double q = 0.0;
                for (int r = 0; r < p; ++r) {
                    double e = yr[r];
                    isn = isn || isnan(e);
                    for (int l = 0; l < L; ++l) e = fma(-__ldg(U + (size_t)r * L + l), wl[l], e);
                    q = fma(e, e, q);
                }
                if (isn) note_nan_row(nan_info, nan_rows, nan_cap, n * T + t0 + row0 + 8 * rb);
                rho_sum += sqrt(q);
            }
        }
    }
    if (rho_part) {
        const double s = block_sum<MT>(rho_sum, red);
        if (tid == 0) rho_part[n * tiles + tile] = s;
    }
}

// Missing observations (moihgp.h:167-178): every listed row gets  u = S^-1/2 (U0'U0)^-1 U0' y_obs  (and w unscaled).
// grid-stride over the list, one CTA per row, scratch in dynamic shared memory.
__global__ void __launch_bounds__(128) k_nan_fix(const double* __restrict__ Y, const double* __restrict__ U, const double* __restrict__ S,
                                                int p, int L, long long T, const int* __restrict__ nan_info,
                                                const long long* __restrict__ nan_rows, long long nan_cap,
                                                double* __restrict__ u, double* __restrict__ w) {
    extern __shared__ double sc[];
    long long count = nan_info[1];
    if (count > nan_cap) count = nan_cap;
    for (long long i = blockIdx.x; i < count; i += gridDim.x) {
        const long long row = nan_rows[i];
        const long long n = row / T, t = row - n * T;
        const double* yr = Y + (size_t)row * p;
        ls_solve_coop(p, L, [&](int r) { return yr[r]; }, [&](int r, int l) { return __ldg(U + (size_t)r * L + l); },
                      sc, sc + L * L, sc + 2 * L * L, sc + 2 * L * L + L, reinterpret_cast<int*>(sc + 2 * L * L + 2 * L),
                      (int)threadIdx.x, (int)blockDim.x, [] { __syncthreads(); });
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            const double z = sc[2 * L * L + L + l];
            const size_t o = ((size_t)n * L + l) * T + t;
            u[o] = z * (1.0 / sqrt(S[l]));
            if (w) w[o] = z;
        }
        __syncthreads();
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

template <int NB>
cudaError_t run_project_mma(const double* Y, const double* U, const double* S, int p, int L, long long N, long long T, double* u,
                            double* w, double* yl, double* rho_part, int* nan_info, long long* nan_rows, long long nan_cap,
                            cudaStream_t stream) {
    using SMC = MmaSmem<NB>;
    static bool attr_done[64] = {};          // per device: function attributes belong to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaFuncSetAttribute(k_project_mma<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMC::BYTES);
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const unsigned grid = (unsigned)(((T + PT - 1) / PT) * N);
    k_project_mma<NB><<<grid, MT, SMC::BYTES, stream>>>(Y, U, S, p, L, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap);
    return cudaGetLastError();
}

}  // namespace

size_t project_tiles(long long T) { return (size_t)((T + PT - 1) / PT); }

cudaError_t launch_project(const double* Y, const double* U, const double* S, int p, int L, long long N, long long T,
                           double* u, double* w, double* yl, double* rho_part, int* nan_info, long long* nan_rows, long long nan_cap,
                           cudaStream_t stream) {
    const bool aligned = (reinterpret_cast<size_t>(Y) & 15) == 0;
    cudaError_t e;
    if (p % 2 == 0 && aligned && L <= 64 && p >= 8) {
        if (L <= 8) e = run_project_mma<1>(Y, U, S, p, L, N, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap, stream);
        else if (L <= 16) e = run_project_mma<2>(Y, U, S, p, L, N, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap, stream);
        else if (L <= 32) e = run_project_mma<4>(Y, U, S, p, L, N, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap, stream);
        else e = run_project_mma<8>(Y, U, S, p, L, N, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap, stream);
    } else {
        const size_t smem = sizeof(double) * ((size_t)PT * (PC + 1) + (size_t)PC * L + (size_t)PT * (L + 1));
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const unsigned grid = (unsigned)(((T + PT - 1) / PT) * N);
        k_project<<<grid, PT, smem, stream>>>(Y, U, S, p, L, T, u, w, yl, rho_part, nan_info, nan_rows, nan_cap);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
    // missing-data rows (usually none: the kernel then exits at once)
    const size_t sc = sizeof(double) * (size_t)ls_scratch_doubles(L);
    if (sc > 48 * 1024) cudaFuncSetAttribute(k_nan_fix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc);
    k_nan_fix<<<592, 128, sc, stream>>>(Y, U, S, p, L, T, nan_info, nan_rows, nan_cap, u, w);
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
