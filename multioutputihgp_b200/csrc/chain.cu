// Many-chains path: one thread per (sequence, latent) chain, sequential in time (sm_100a).
//
// When there are tens of thousands of independent (sequence, latent) recurrences (BASELINE config 3:
// 4096 x 8), time-parallel scanning only adds redundant flops; here every recurrence is evaluated
// exactly once, literally as the reference's loop does, and the kernels are pure HBM streams.
//
// Replaces the same reference code as project.cu + scan.cu:
//   moihgp.h:159-182, :499-501   projection and residual norm            (phase 1 of k_filter_chain)
//   ihgp.h:81-93, :204-209       IHGP::step / IHGP::negLogLikelihood     (phase 2 of k_filter_chain)
//   moihgp.h:614-688             MOIHGP::negLogLikelihood(x, y)          (in-warp reduction, k_filter_chain)
//   ihgp.h:108-113               IHGP::backwardSmoother recursion        (k_smooth_chain; literal and RTS forms)
//
// k_filter_chain<P, L, D>: one WARP owns NS = 32 / L sequences for all T steps, in rounds of L steps.
//   load   : the [NS][L rows][P] tile of Y (NS contiguous runs of L*P*8 bytes) is staged by cp.async into a
//            ring of shared-memory stages; the 16-byte chunks of each row are XOR-swizzled with the row index so
//            that "one row per lane" reads are bank-conflict free.
//   phase 1: lane (s, j) projects row j of sequence s onto all L latents (U^T y, with U in the constant bank),
//            forms the residual norm || y - U U^T y ||, and drops u = S^-1/2 U^T y into a small exchange tile.
//   phase 2: lane (s, l) picks latent l of the L steps out of the exchange tile and runs the recurrence
//            x+ = AKHA x + K u, accumulating the innovation likelihood; states go to a staging tile.
//   store  : the [NS][L rows][L*D] staging tile leaves as NS contiguous runs with 16-byte stores.
// k_smooth_chain<L, D, MODE>: the same lane mapping streaming backwards over the stored X.
#include <cuda_runtime.h>
#include <math.h>
#include <cmath>
#include "moihgp_device.cuh"
#include "launch.h"

namespace moihgp {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STAGES = 4;

template <int P, int L>
struct ProjConsts {          // passed by value: lives in the constant bank, uniform operands of the phase-1 FMAs
    double U[P][L];
    double rs[L];            // S^-1/2
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int P, int L, int D>
struct FilterSmem {
    static constexpr int NS = 32 / L;
    static constexpr int ROWB = P * 8;                 // bytes per row of Y
    static constexpr int CH = ROWB / 16;               // 16-byte chunks per row
    static constexpr int TILE = NS * L * ROWB;         // bytes per stage
    static constexpr int XROW = L + 2;                 // exchange tile row pitch (doubles): conflict-free STS.128
    static constexpr int XSEQ = L * XROW + 8;          // per-sequence pitch (doubles)
    static constexpr int OSEQ = L * L * D + 4;         // staging tile per-sequence pitch (doubles), 16B-aligned
    static constexpr int BYTES = STAGES * TILE + NS * XSEQ * 8 + NS * OSEQ * 8;
};

// grid: ceil(N / NS) CTAs of ONE warp.
template <int P, int L, int D>
__global__ void __launch_bounds__(32) k_filter_chain(const double* __restrict__ Y, const __grid_constant__ ProjConsts<P, L> pc,
                                                    const LatentConsts* __restrict__ consts, double sigma, double nll_const,
                                                    long long N, long long T, const double* __restrict__ x0,
                                                    double* __restrict__ X, double* __restrict__ nll, double* __restrict__ xT) {
    using SM = FilterSmem<P, L, D>;
    constexpr int NS = SM::NS, CH = SM::CH;
    static_assert(P % 2 == 0 && CH >= 1 && (CH & (CH - 1)) == 0, "P*8 bytes must be a power-of-two number of 16B chunks");
    static_assert(32 % L == 0, "L must divide the warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* ytile = smem_raw;                                           // [STAGES][NS][L][P] swizzled
    double* xch = reinterpret_cast<double*>(smem_raw + STAGES * SM::TILE);     // [NS][L][XROW]
    double* ost = xch + NS * SM::XSEQ;                                         // [NS][L][L*D]
    const int lane = threadIdx.x;
    const int s = lane / L, j = lane % L;                                      // phase 1: (sequence, row); phase 2: (sequence, latent)
    const long long n0 = (long long)blockIdx.x * NS;
    const long long n = n0 + s;
    const bool seq_ok = n < N;
    const long long rounds = (T + L - 1) / L;

    // ---- per-lane latent constants (phase 2) ---------------------------------------------------
    const LatentConsts* lc = consts + j;
    double M[D * D], K[D], HA[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        K[a] = __ldg(&lc->K[a]);
        HA[a] = __ldg(&lc->HA[a]);
#pragma unroll
        for (int b = 0; b < D; ++b) M[a * D + b] = __ldg(&lc->AKHA[a * 3 + b]);
    }
    double x[D];
#pragma unroll
    for (int a = 0; a < D; ++a) x[a] = (x0 && seq_ok) ? x0[((size_t)n * L + j) * D + a] : 0.0;
    double rho_acc = 0.0, vsq_acc = 0.0;

    // ---- cp.async producer: tile of round r into stage r % STAGES --------------------------------
    constexpr int PIECES = NS * L * CH;            // 16-byte pieces per tile
    auto issue = [&](long long r) {
        if (r < rounds) {
            unsigned char* st = ytile + (size_t)(r % STAGES) * SM::TILE;
            const long long t0 = r * L;
#pragma unroll
            for (int k = 0; k < PIECES / 32; ++k) {
                const int q = lane + 32 * k;
                const int qs = q / (L * CH), qj = (q / CH) % L, qc = q % CH;
                if (n0 + qs < N && t0 + qj < T)
                    cp_async16(st + (size_t)(qs * L + qj) * SM::ROWB + ((qc ^ (qj % CH)) * 16),
                               reinterpret_cast<const unsigned char*>(Y) + (((size_t)(n0 + qs) * T + t0 + qj) * P) * 8 + qc * 16);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int r = 0; r < STAGES - 1; ++r) issue(r);

    for (long long r = 0; r < rounds; ++r) {
        issue(r + STAGES - 1);
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        const long long t0 = r * L;
        // ---- phase 1: project row (s, j) -----------------------------------------------------------
        {
            const unsigned char* row = ytile + (size_t)(r % STAGES) * SM::TILE + (size_t)(s * L + j) * SM::ROWB;
            double y[P];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const double2 v = *reinterpret_cast<const double2*>(row + ((c ^ (j % CH)) * 16));
                y[2 * c] = v.x;
                y[2 * c + 1] = v.y;
            }
            const bool row_ok = seq_ok && (t0 + j < T);
            if (!row_ok) {
#pragma unroll
                for (int c = 0; c < P; ++c) y[c] = 0.0;
            }
            double w[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                double acc = pc.U[0][l] * y[0];
#pragma unroll
                for (int c = 1; c < P; ++c) acc = fma(pc.U[c][l], y[c], acc);     // U' y            moihgp.h:181
                w[l] = acc;
            }
            double q = 0.0;
#pragma unroll
            for (int c = 0; c < P; ++c) {
                double e = y[c];
#pragma unroll
                for (int l = 0; l < L; ++l) e = fma(-pc.U[c][l], w[l], e);        // (I - U U') y    moihgp.h:651
                q = fma(e, e, q);
            }
            rho_acc += sqrt(q);                                                   // norm, not squared (Q9)
            double* xr = xch + s * SM::XSEQ + j * SM::XROW;
#pragma unroll
            for (int l = 0; l < L; l += 2) *reinterpret_cast<double2*>(xr + l) = make_double2(w[l] * pc.rs[l], w[l + 1] * pc.rs[l + 1]);
        }
        __syncwarp();
        // ---- phase 2: latent (s, j) over the L steps of the round --------------------------------------
        {
            const double* xc = xch + s * SM::XSEQ + j;
            double* oc = ost + s * SM::OSEQ + j * D;
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const double u = xc[i * SM::XROW];
                double hax = HA[0] * x[0];
#pragma unroll
                for (int a = 1; a < D; ++a) hax = fma(HA[a], x[a], hax);
                const double v = u - hax;                                         // ihgp.h:206
                double xn[D];
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    double acc = M[a * D] * x[0];
#pragma unroll
                    for (int b = 1; b < D; ++b) acc = fma(M[a * D + b], x[b], acc);
                    xn[a] = fma(K[a], u, acc);                                    // ihgp.h:90
                }
                if (t0 + i < T) {
                    vsq_acc = fma(v, v, vsq_acc);
#pragma unroll
                    for (int a = 0; a < D; ++a) x[a] = xn[a];
                }
#pragma unroll
                for (int a = 0; a < D; ++a) oc[i * (L * D) + a] = x[a];
            }
        }
        __syncwarp();
        // ---- store: NS contiguous runs of L rows ------------------------------------------------------
        if (X) {
            constexpr int RUN16 = L * L * D / 2;                                  // 16-byte pieces per sequence-round
            const long long rows_left = T - t0;
            const int valid16 = (int)(rows_left >= L ? RUN16 : rows_left * (L * D / 2));
#pragma unroll
            for (int k = 0; k < (NS * RUN16 + 31) / 32; ++k) {
                const int q = lane + 32 * k;
                const int qs = q / RUN16, qo = q % RUN16;
                if (q < NS * RUN16 && n0 + qs < N && qo < valid16) {
                    const double2 v = *reinterpret_cast<const double2*>(ost + qs * SM::OSEQ + 2 * qo);
                    *reinterpret_cast<double2*>(X + ((size_t)(n0 + qs) * T + t0) * (L * D) + 2 * qo) = v;
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    // ---- final state and NLL of each sequence ---------------------------------------------------------
    if (xT && seq_ok) {
#pragma unroll
        for (int a = 0; a < D; ++a) xT[((size_t)n * L + j) * D + a] = x[a];
    }
    if (nll) {
        // sum_t 1/2 rho_t / sigma  +  sum_l 1/2 sum_t v^2 / S_l          moihgp.h:653, ihgp.h:207
        double part = 0.5 * rho_acc / sigma + 0.5 * vsq_acc / __ldg(&lc->S);
#pragma unroll
        for (int o = 1; o < L; o <<= 1) part += __shfl_xor_sync(FULL, part, o);
        if (j == 0 && seq_ok) nll[n] = part + nll_const;
    }
}

template <int L, int D>
struct SmoothSmem {
    static constexpr int NS = 32 / L;
    static constexpr int RUN = L * L * D;              // doubles per sequence-round
    static constexpr int OSEQ = RUN + 4;
    static constexpr int BYTES = STAGES * NS * OSEQ * 8;
};

// Backward sweep over the stored filtered states.  MODE 0: reference_literal (ihgp.h:108-113, Q3)
//   Xs[T-1] = X[T-1];  Xs[j] = X[j+1] + G Xs[j+1] - A X[j+1]
// MODE 1: rts_correct   Xs[j] = X[j] + G (Xs[j+1] - A X[j]).
// grid: ceil(N / NS) CTAs of one warp; rounds of L steps from the end of the sequence.
template <int L, int D, int MODE>
__global__ void __launch_bounds__(32) k_smooth_chain(const double* __restrict__ X, const LatentConsts* __restrict__ consts,
                                                    long long N, long long T, double* __restrict__ Xs) {
    using SM = SmoothSmem<L, D>;
    constexpr int NS = SM::NS, RUN = SM::RUN, RUN16 = RUN / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);                       // [STAGES][NS][OSEQ]
    const int lane = threadIdx.x;
    const int s = lane / L, j = lane % L;
    const long long n0 = (long long)blockIdx.x * NS;
    const bool seq_ok = n0 + s < N;
    const long long rounds = (T + L - 1) / L;
    const LatentConsts* lc = consts + j;
    double G[D * D], A[D * D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) { G[a * D + b] = __ldg(&lc->G[MODE][a * 3 + b]); A[a * D + b] = __ldg(&lc->A[a * 3 + b]); }

    // round index k counts from the END: it covers steps t0 = (rounds - 1 - k) * L ...
    auto issue = [&](long long k) {
        if (k < rounds) {
            double* st = tiles + (size_t)(k % STAGES) * NS * SM::OSEQ;
            const long long t0 = (rounds - 1 - k) * L;
            const long long rows_left = T - t0;
            const int valid16 = (int)(rows_left >= L ? RUN16 : rows_left * (L * D / 2));
#pragma unroll
            for (int m = 0; m < (NS * RUN16 + 31) / 32; ++m) {
                const int q = lane + 32 * m;
                const int qs = q / RUN16, qo = q % RUN16;
                if (q < NS * RUN16 && n0 + qs < N && qo < valid16)
                    cp_async16(st + qs * SM::OSEQ + 2 * qo, X + ((size_t)(n0 + qs) * T + t0) * (L * D) + 2 * qo);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < STAGES - 1; ++k) issue(k);

    double xs[D], xnext[D];                       // Xs[j+1] and X[j+1]
#pragma unroll
    for (int a = 0; a < D; ++a) { xs[a] = 0.0; xnext[a] = 0.0; }
    for (long long k = 0; k < rounds; ++k) {
        issue(k + STAGES - 1);
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        double* st = tiles + (size_t)(k % STAGES) * NS * SM::OSEQ + s * SM::OSEQ + j * D;
        const long long t0 = (rounds - 1 - k) * L;
#pragma unroll
        for (int i = L - 1; i >= 0; --i) {
            const long long t = t0 + i;
            if (t < T) {
                double xx[D], out[D];
#pragma unroll
                for (int a = 0; a < D; ++a) xx[a] = st[i * (L * D) + a];
                if (t == T - 1) {
#pragma unroll
                    for (int a = 0; a < D; ++a) out[a] = xx[a];                    // ihgp.h:108
                } else if (MODE == 0) {
#pragma unroll
                    for (int a = 0; a < D; ++a) {                                  // ihgp.h:111
                        double g = 0.0, aa = 0.0;
#pragma unroll
                        for (int b = 0; b < D; ++b) { g = fma(G[a * D + b], xs[b], g); aa = fma(A[a * D + b], xnext[b], aa); }
                        out[a] = xnext[a] + g - aa;
                    }
                } else {
                    double rr[D];
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        double aa = 0.0;
#pragma unroll
                        for (int b = 0; b < D; ++b) aa = fma(A[a * D + b], xx[b], aa);
                        rr[a] = xs[a] - aa;
                    }
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        double g = 0.0;
#pragma unroll
                        for (int b = 0; b < D; ++b) g = fma(G[a * D + b], rr[b], g);
                        out[a] = xx[a] + g;
                    }
                }
#pragma unroll
                for (int a = 0; a < D; ++a) { xs[a] = out[a]; xnext[a] = xx[a]; st[i * (L * D) + a] = out[a]; }
            }
        }
        __syncwarp();
        {
            double* stw = tiles + (size_t)(k % STAGES) * NS * SM::OSEQ;
            const long long rows_left = T - t0;
            const int valid16 = (int)(rows_left >= L ? RUN16 : rows_left * (L * D / 2));
#pragma unroll
            for (int m = 0; m < (NS * RUN16 + 31) / 32; ++m) {
                const int q = lane + 32 * m;
                const int qs = q / RUN16, qo = q % RUN16;
                if (q < NS * RUN16 && n0 + qs < N && qo < valid16) {
                    const double2 v = *reinterpret_cast<const double2*>(stw + qs * SM::OSEQ + 2 * qo);
                    *reinterpret_cast<double2*>(Xs + ((size_t)(n0 + qs) * T + t0) * (L * D) + 2 * qo) = v;
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    (void)seq_ok;
}

template <int P, int L, int D>
cudaError_t run_chain(const ChainArgs& a, cudaStream_t st) {
    using FS = FilterSmem<P, L, D>;
    using SS = SmoothSmem<L, D>;
    ProjConsts<P, L> pc;
    for (int r = 0; r < P; ++r) for (int l = 0; l < L; ++l) pc.U[r][l] = a.U_host[(size_t)r * L + l];
    for (int l = 0; l < L; ++l) pc.rs[l] = 1.0 / std::sqrt(a.S_host[l]);
    const unsigned grid = (unsigned)((a.N + FS::NS - 1) / FS::NS);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_filter_chain<P, L, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS::BYTES);
        cudaFuncSetAttribute(k_smooth_chain<L, D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS::BYTES);
        cudaFuncSetAttribute(k_smooth_chain<L, D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS::BYTES);
        attr_done = true;
    }
    k_filter_chain<P, L, D><<<grid, 32, FS::BYTES, st>>>(a.Y, pc, a.consts, a.sigma, a.nll_const, a.N, a.T, a.x0, a.X, a.nll, a.xT);
    mark(a.mk, "k_filter_chain");
    if (a.Xs) {
        if (a.mode == 0) k_smooth_chain<L, D, 0><<<grid, 32, SS::BYTES, st>>>(a.X, a.consts, a.N, a.T, a.Xs);
        else k_smooth_chain<L, D, 1><<<grid, 32, SS::BYTES, st>>>(a.X, a.consts, a.N, a.T, a.Xs);
        mark(a.mk, "k_smooth_chain");
    }
    return cudaGetLastError();
}

}  // namespace

bool chain_supported(int p, int L, int dim) {
    return (p == 16 && L == 8) || (p == 8 && L == 4);
}

cudaError_t launch_chain(int p, int L, int dim, const ChainArgs& a, cudaStream_t st) {
    if (p == 16 && L == 8) return dim == 3 ? run_chain<16, 8, 3>(a, st) : run_chain<16, 8, 2>(a, st);
    if (p == 8 && L == 4) return dim == 3 ? run_chain<8, 4, 3>(a, st) : run_chain<8, 4, 2>(a, st);
    return cudaErrorInvalidValue;
}

}  // namespace moihgp
