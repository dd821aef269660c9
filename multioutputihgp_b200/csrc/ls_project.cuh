// Missing-observation (NaN) projection of ONE observation (sm_100a), cooperative over a group of threads.
//
// Reference: moihgp.h:150-178 (and the identical blocks at :231-263, :306-336, :462-498, :616-648): with U0 the rows of U
// whose output is observed,
//     Ty = diag(S^-1/2) (U0' U0).ldlt().solve(U0' y_obs).
// Eigen's LDLT pivots symmetrically on the largest remaining diagonal entry and its solve() maps an (almost) zero pivot
// to 0 - so an all-NaN observation, whose U0'U0 is the zero matrix, gives Ty = 0.  The routine below follows the same
// factorisation.
#pragma once
#include <cfloat>
#include <math.h>

namespace moihgp {

// Solve (U0'U0) z = U0' y_obs for one row.  `tid`/`nt`: this thread's index in the cooperating group and the group size;
// `sync()`: the group's barrier (__syncwarp or __syncthreads).  y(r): the observation (NaN = missing); Uat(r, l): U[r][l].
// Scratch (shared memory, doubles): M[L*L], Lm[L*L], D[L], v[L], and ints perm[L].  Result z[0..L) left in v.
template <typename YF, typename UF, typename SyncF>
__device__ __forceinline__ void ls_solve_coop(int p, int L, YF y, UF Uat, double* M, double* Lm, double* D, double* v, int* perm,
                                              int tid, int nt, SyncF sync) {
    // G = U0'U0, b = U0'y_obs (sums over the observed rows in ascending order)
    for (int e = tid; e < L * L + L; e += nt) {
        if (e < L * L) {
            const int a = e / L, c = e - a * L;
            double s = 0.0;
            for (int r = 0; r < p; ++r) { const double yr = y(r); if (yr == yr) s += Uat(r, a) * Uat(r, c); }
            M[e] = s;
            Lm[e] = a == c ? 1.0 : 0.0;
        } else {
            const int a = e - L * L;
            double s = 0.0;
            for (int r = 0; r < p; ++r) { const double yr = y(r); if (yr == yr) s += Uat(r, a) * yr; }
            v[a] = s;
            perm[a] = a;
        }
    }
    sync();
    for (int k = 0; k < L; ++k) {
        // pivot: largest |diagonal| of the trailing block (first one wins ties)
        int piv = k;
        for (int i = k + 1; i < L; ++i) if (fabs(M[i * L + i]) > fabs(M[piv * L + piv])) piv = i;   // every thread computes the same piv
        sync();
        if (piv != k) {
            for (int j = tid; j < L; j += nt) { const double t = M[k * L + j]; M[k * L + j] = M[piv * L + j]; M[piv * L + j] = t; }
            sync();
            for (int i = tid; i < L; i += nt) { const double t = M[i * L + k]; M[i * L + k] = M[i * L + piv]; M[i * L + piv] = t; }
            for (int j = tid; j < k; j += nt) { const double t = Lm[k * L + j]; Lm[k * L + j] = Lm[piv * L + j]; Lm[piv * L + j] = t; }
            if (tid == 0) { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
            sync();
        }
        const double dk = M[k * L + k];
        if (tid == 0) D[k] = dk;
        if (dk != 0.0) {
            for (int i = k + 1 + tid; i < L; i += nt) Lm[i * L + k] = M[i * L + k] / dk;
            sync();
            const int m = L - k - 1;
            for (int e = tid; e < m * m; e += nt) {
                const int i = k + 1 + e / m, j = k + 1 + e % m;
                if (i >= j) {
                    const double nv = M[i * L + j] - Lm[i * L + k] * dk * Lm[j * L + k];
                    M[i * L + j] = nv;
                    if (i != j) M[j * L + i] = nv;
                }
            }
        }
        sync();
    }
    // forward / diagonal / backward substitution on the permuted right-hand side (tiny: one thread)
    if (tid == 0) {
        double* yv = M;                      // M is dead: reuse its first 2 L entries
        double* bv = M + L;
        for (int i = 0; i < L; ++i) bv[i] = v[perm[i]];
        for (int i = 0; i < L; ++i) { double s = bv[i]; for (int j = 0; j < i; ++j) s -= Lm[i * L + j] * yv[j]; yv[i] = s; }
        for (int i = 0; i < L; ++i) yv[i] = fabs(D[i]) > DBL_MIN ? yv[i] / D[i] : 0.0;
        for (int i = L - 1; i >= 0; --i) { double s = yv[i]; for (int j = i + 1; j < L; ++j) s -= Lm[j * L + i] * yv[j]; yv[i] = s; }
        for (int i = 0; i < L; ++i) v[perm[i]] = yv[i];
    }
    sync();
}

// doubles of shared-memory scratch ls_solve_coop needs (the int perm[L] array is carved from the tail)
__host__ __device__ constexpr int ls_scratch_doubles(int L) { return 2 * L * L + 2 * L + (L + 1) / 2 + 1; }

}  // namespace moihgp
