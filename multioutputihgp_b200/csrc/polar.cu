// K-polar: the polar factor U = W V' of the p x L block of the parameter vector, on the device (sm_100a).
//
// Replaces MOIHGP::update's  U = svd.matrixU() * svd.matrixV().transpose()  (moihgp.h:431-447; JacobiSVD / BDCSVD) for
// large p*L, where the host Jacobi costs milliseconds per objective evaluation.  One-sided Jacobi on
// A V = W diag(s): columns live contiguously in shared memory (Wt[j][r], Vt[j][r]); each sweep is a round-robin
// tournament of Lp - 1 rounds with Lp / 2 disjoint column pairs, one warp per pair per round.  The polar factor is
// unique, so the rotation order only moves the result by rounding (~1e-16).
#include <cuda_runtime.h>
#include <math.h>
#include "launch.h"

namespace moihgp {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// one CTA; dynamic shared memory: Wt[Lp][p], Vt[Lp][Lp], norms[Lp]
__global__ void __launch_bounds__(1024) k_polar(const double* __restrict__ A /*[p][L] row-major*/, int p, int L, double* __restrict__ U) {
    extern __shared__ double sm[];
    __shared__ double off_max;
    const int Lp = (L + 1) & ~1;
    double* Wt = sm;
    double* Vt = Wt + (size_t)Lp * p;
    double* nrm = Vt + (size_t)Lp * Lp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < Lp * p; i += blockDim.x) {
        const int c = i / p, r = i - c * p;
        Wt[i] = c < L ? A[(size_t)r * L + c] : 0.0;
    }
    for (int i = tid; i < Lp * Lp; i += blockDim.x) Vt[i] = (i / Lp) == (i % Lp) ? 1.0 : 0.0;
    __syncthreads();
    const int n1 = Lp - 1;
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (tid == 0) off_max = 0.0;
        __syncthreads();
        double my_off = 0.0;
        for (int rnd = 0; rnd < n1; ++rnd) {
            for (int pr = warp; pr < Lp / 2; pr += nwarps) {
                int i, j;
                if (pr == 0) { i = n1; j = rnd; }
                else {                                   // (rnd +- pr) mod n1 with 0 <= rnd, pr < n1
                    i = rnd + pr; if (i >= n1) i -= n1;
                    j = rnd - pr; if (j < 0) j += n1;
                }
                if (i > j) { const int t = i; i = j; j = t; }
                double* wi = Wt + (size_t)i * p;
                double* wj = Wt + (size_t)j * p;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = lane; r < p; r += 32) { const double x = wi[r], y = wj[r]; al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga); }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                // relative size of the off-diagonal entry, |ga| / sqrt(al be), compared through squares (no sqrt / division on
                // this latency chain): >= 1e-15 keeps the sweeps going, <= 3e-17 is not worth a rotation
                const double g2 = ga * ga, prod = al * be;
                if (g2 > 1e-30 * prod) my_off = 1.0;
                if (g2 > 1e-33 * prod) {
                    // Jacobi angle, |theta| <= pi/4, tan(2 theta) = 2 ga / (be - al), without a division on the chain:
                    // cos(2 theta) = |d| / h, h = hypot(d, 2 ga);  c = sqrt((1 + cos 2theta) / 2);  s = sin(2 theta) / (2 c)
                    const double d = be - al;
                    const double rh = rsqrt(fma(d, d, 4.0 * g2));
                    const double x = fma(0.5 * fabs(d), rh, 0.5);
                    const double ic = rsqrt(x);
                    const double c = x * ic, s = copysign(ga * rh * ic, d * ga);
                    for (int r = lane; r < p; r += 32) { const double x = wi[r], y = wj[r]; wi[r] = c * x - s * y; wj[r] = s * x + c * y; }
                    double* vi = Vt + (size_t)i * Lp;
                    double* vj = Vt + (size_t)j * Lp;
                    for (int r = lane; r < Lp; r += 32) { const double x = vi[r], y = vj[r]; vi[r] = c * x - s * y; vj[r] = s * x + c * y; }
                }
            }
            __syncthreads();
        }
        if (lane == 0 && my_off > 0.0) off_max = 1.0;               // some pair was still above 1e-15 in this sweep
        __syncthreads();
        const double off = off_max;
        __syncthreads();
        if (off == 0.0) break;
    }
    for (int j = warp; j < Lp; j += nwarps) {
        const double* wj = Wt + (size_t)j * p;
        double q = 0.0;
        for (int r = lane; r < p; r += 32) q = fma(wj[r], wj[r], q);
        q = warp_sum(q);
        if (lane == 0) nrm[j] = q > 0.0 ? 1.0 / sqrt(q) : 0.0;       // 1 / singular value (0 for a zero column)
    }
    __syncthreads();
    for (int i = tid; i < L * p; i += blockDim.x) Wt[i] *= nrm[i / p];   // W <- W diag(1/s): orthonormal columns
    __syncthreads();
    for (int i = tid; i < p * L; i += blockDim.x) {
        const int r = i / L, c = i - r * L;
        double s = 0.0;
        for (int k = 0; k < L; ++k) s = fma(Wt[(size_t)k * p + r], Vt[(size_t)k * Lp + c], s);
        U[i] = s;
    }
}

}  // namespace

size_t polar_smem_bytes(int p, int L) {
    const size_t Lp = (size_t)((L + 1) & ~1);
    return sizeof(double) * (Lp * p + Lp * Lp + Lp);
}

cudaError_t launch_polar(const double* A, int p, int L, double* U, cudaStream_t st) {
    const size_t smem = polar_smem_bytes(p, L);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static bool attr_done[64] = {};          // per device: function attributes belong to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaFuncSetAttribute(k_polar, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int Lp = (L + 1) & ~1;
    int warps = Lp / 2;
    if (warps > 32) warps = 32;
    if (warps < 1) warps = 1;
    k_polar<<<1, 32 * warps, smem, st>>>(A, p, L, U);
    return cudaGetLastError();
}

}  // namespace moihgp
