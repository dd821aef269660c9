// =====================================================================================
// TEST INFRASTRUCTURE ONLY - the CPU oracle for the MOIHGP hot path.
//
// An Eigen-free C++ restatement of the reference's algorithm (lim271/MultiOutputIHGP),
// following the reference's operation order and ALL of its arithmetic quirks (SURVEY.md
// section 9).  Every function cites the reference file:line it follows (paths relative to
// /root/reference/moihgp/include unless stated).
//
// Pinning: the reference ships no tests or golden vectors.  This restatement is pinned
// against the UNMODIFIED reference sources compiled in this container against an Eigen-API
// shim (oracle/_ref, see oracle/Makefile and oracle/eigen_shim/Eigen/Core) - see
// tests/test_oracle_vs_ref.py (runs where oracle/_ref exists) and the fixtures under
// tests/golden/ generated from that build by oracle/gen_golden.py.  The shim supplies the
// arithmetic of expm / SVD / LDLT (Eigen itself is absent from the image); everything else
// - operation order, indexing, the quirks - is the reference's own code.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library, and only as the checker / the timed CPU baseline.  The product
// (multioutputihgp_b200/) never links, loads or calls it.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------ small dense matrix
struct SM {  // row-major, up to 6x6
    int r, c;
    double a[36];
    SM() : r(0), c(0) { std::memset(a, 0, sizeof(a)); }
    SM(int r_, int c_) : r(r_), c(c_) { std::memset(a, 0, sizeof(a)); }
    double& operator()(int i, int j) { return a[i * c + j]; }
    double operator()(int i, int j) const { return a[i * c + j]; }
};
static SM eye(int n) { SM m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
static SM mul(const SM& x, const SM& y) {
    SM m(x.r, y.c);
    for (int i = 0; i < x.r; ++i) for (int j = 0; j < y.c; ++j) { double s = 0.0; for (int k = 0; k < x.c; ++k) s += x(i, k) * y(k, j); m(i, j) = s; }
    return m;
}
static SM add(const SM& x, const SM& y) { SM m(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) m.a[i] = x.a[i] + y.a[i]; return m; }
static SM sub(const SM& x, const SM& y) { SM m(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) m.a[i] = x.a[i] - y.a[i]; return m; }
static SM scale(const SM& x, double s) { SM m(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) m.a[i] = x.a[i] * s; return m; }
static SM divs(const SM& x, double s) { SM m(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) m.a[i] = x.a[i] / s; return m; }
static SM neg(const SM& x) { return scale(x, -1.0); }
static SM tr(const SM& x) { SM m(x.c, x.r); for (int i = 0; i < x.r; ++i) for (int j = 0; j < x.c; ++j) m(j, i) = x(i, j); return m; }
static bool is_zero(const SM& x) { for (int i = 0; i < x.r * x.c; ++i) if (!(x.a[i] == 0.0)) return false; return true; }
static double max_coeff(const SM& x) { double m = x.a[0]; for (int i = 1; i < x.r * x.c; ++i) m = x.a[i] > m ? x.a[i] : m; return m; }

// Solve A X = B (partial-pivot Gaussian elimination), A n x n, n <= 6.
static SM solve_lu(SM A, SM B) {
    const int n = A.r;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(A(i, k)) > std::fabs(A(piv, k))) piv = i;
        if (piv != k) { for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j)); for (int j = 0; j < B.c; ++j) std::swap(B(k, j), B(piv, j)); }
        for (int i = k + 1; i < n; ++i) {
            const double f = A(i, k) / A(k, k);
            for (int j = k; j < n; ++j) A(i, j) -= f * A(k, j);
            for (int j = 0; j < B.c; ++j) B(i, j) -= f * B(k, j);
        }
    }
    SM X(n, B.c);
    for (int j = 0; j < B.c; ++j) for (int i = n - 1; i >= 0; --i) {
        double s = B(i, j);
        for (int k = i + 1; k < n; ++k) s -= A(i, k) * X(k, j);
        X(i, j) = s / A(i, i);
    }
    return X;
}

// Matrix exponential.  The reference calls Eigen's (_dt * F).exp() (ihgp.h:120,167), which is
// Higham's 2005 scaling-and-squaring Pade algorithm; the published [13/13] variant is restated
// here (always degree 13, scaled so that ||A||_1 <= 5.37).  Any accurate expm agrees ~1e-15.
static SM expm(const SM& Ain) {
    const int n = Ain.r;
    double l1 = 0.0;
    for (int j = 0; j < n; ++j) { double s = 0.0; for (int i = 0; i < n; ++i) s += std::fabs(Ain(i, j)); l1 = std::max(l1, s); }
    int sq = 0;
    if (l1 > 5.371920351148152) { int e = 0; std::frexp(l1 / 5.371920351148152, &e); sq = e > 0 ? e : 0; }
    const SM A = scale(Ain, std::ldexp(1.0, -sq));
    static const double b[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                               129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                               40840800., 960960., 16380., 182., 1.};
    const SM I = eye(n), A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2);
    SM U = add(add(scale(A6, b[13]), scale(A4, b[11])), scale(A2, b[9]));
    U = add(add(add(add(mul(A6, U), scale(A6, b[7])), scale(A4, b[5])), scale(A2, b[3])), scale(I, b[1]));
    U = mul(A, U);
    SM V = add(add(scale(A6, b[12]), scale(A4, b[10])), scale(A2, b[8]));
    V = add(add(add(add(mul(A6, V), scale(A6, b[6])), scale(A4, b[4])), scale(A2, b[2])), scale(I, b[0]));
    SM R = solve_lu(sub(V, U), add(V, U));
    for (int i = 0; i < sq; ++i) R = mul(R, R);
    return R;
}

// ------------------------------------------------------------------ utils/dare.h
static const double dare_tol = 1e-8;          // dare.h:7
static const unsigned dare_maxiter = 100;     // dare.h:8

// dare.h:10-33.  Fixed-point Riccati iteration from P0 = Q; signed-max stop test (Q1).
// Returns the iteration count in *iters (1-based; 100 when the limit is hit).
static bool DARE(const SM& Ad, const SM& Bd, const SM& Q, const SM& R, SM& P, int* iters) {
    P = Q;                                                                      // dare.h:12
    const SM AdT = tr(Ad), BdT = tr(Bd);                                        // dare.h:16-17
    for (unsigned it = 0; it < dare_maxiter; ++it) {                            // dare.h:20
        // dare.h:23, evaluated left to right like the C++ expression
        const SM AtP = mul(AdT, P);
        const SM T1 = mul(AtP, Ad);
        SM inv(1, 1);
        inv(0, 0) = 1.0 / (add(R, mul(mul(BdT, P), Bd)))(0, 0);
        const SM T2 = mul(mul(mul(mul(mul(AtP, Bd), inv), BdT), P), Ad);
        const SM Pn = add(sub(T1, T2), Q);
        const double diff = std::fabs(max_coeff(sub(Pn, P)));                   // dare.h:25
        P = divs(add(Pn, tr(Pn)), 2.0);                                         // dare.h:26
        if (diff < dare_tol) { if (iters) *iters = int(it) + 1; return true; }  // dare.h:27-30
    }
    if (iters) *iters = int(dare_maxiter);
    return false;                                                               // dare.h:32
}

// dare.h:36-58.  NOT a Lyapunov solve: iterates P <- A'PA - P + Q (Q2).
static bool DLyap(const SM& Ad, const SM& Q, SM& P, int* iters) {
    P = Q;                                                                      // dare.h:38
    const SM AdT = tr(Ad);                                                      // dare.h:42
    for (unsigned it = 0; it < dare_maxiter; ++it) {
        const SM Pn = add(sub(mul(mul(AdT, P), Ad), P), Q);                     // dare.h:48
        const double diff = std::fabs(max_coeff(sub(Pn, P)));                   // dare.h:50
        P = divs(add(Pn, tr(Pn)), 2.0);                                         // dare.h:51
        if (diff < dare_tol) { if (iters) *iters = int(it) + 1; return true; }
    }
    if (iters) *iters = int(dare_maxiter);
    return false;
}

// ------------------------------------------------------------------ state spaces
struct StateSpace {
    int dim;
    SM F, Pinf, H, R, dF[3], dPinf[3], dR[3];
    double params[3];
};

// matern32ss.h:17-64 (constructor + update)
static void matern32(StateSpace& s, const double* p) {
    s.dim = 2;
    s.F = SM(2, 2); s.F(0, 1) = 1.0;                                            // :19-20
    s.Pinf = SM(2, 2);
    s.H = SM(1, 2); s.H(0, 0) = 1.0;                                            // :22-23
    s.R = SM(1, 1);
    for (int k = 0; k < 3; ++k) { s.dF[k] = SM(2, 2); s.dPinf[k] = SM(2, 2); s.dR[k] = SM(1, 1); }
    s.dPinf[0] = eye(2);                                                        // :27
    s.dR[2](0, 0) = 1.0;                                                        // :33
    const double magnitude = p[0], lengthscale = p[1];                          // :42-43
    const double lam = std::sqrt(3) / lengthscale;                              // :44
    const double lam2 = lam * lam;                                              // :45
    const double len3 = 6.0 / (lengthscale * lengthscale * lengthscale);        // :46
    s.F(1, 0) = -lam2; s.F(1, 1) = -2.0 * lam;                                  // :47-48
    s.Pinf(0, 0) = magnitude; s.Pinf(1, 1) = magnitude * lam2;                  // :49-50
    s.R(0, 0) = p[2];                                                           // :51
    s.dF[1](1, 0) = len3; s.dF[1](1, 1) = 2.0 * lam / lengthscale;              // :54-55
    s.dPinf[0](1, 1) = lam2;                                                    // :58
    s.dPinf[1](1, 1) = -magnitude * len3;                                       // :61
    for (int k = 0; k < 3; ++k) s.params[k] = p[k];
}

// matern52ss.h:17-75.  NB lam = sqrt(3)/l in F but sqrt(5)-consistent Pinf/dF (Q4).
static void matern52(StateSpace& s, const double* p) {
    s.dim = 3;
    s.F = SM(3, 3); s.F(0, 1) = 1.0; s.F(1, 2) = 1.0;                           // :19-21
    s.Pinf = SM(3, 3);
    s.H = SM(1, 3); s.H(0, 0) = 1.0;
    s.R = SM(1, 1);
    for (int k = 0; k < 3; ++k) { s.dF[k] = SM(3, 3); s.dPinf[k] = SM(3, 3); s.dR[k] = SM(1, 1); }
    s.dR[2](0, 0) = 1.0;
    const double magnitude = p[0], lengthscale = p[1];
    const double lam = std::sqrt(3.0) / lengthscale;                            // :42
    const double lam2 = lam * lam;
    const double len2 = lengthscale * lengthscale, len3 = len2 * lengthscale, len4 = len2 * len2;
    const double kappa = 5.0 / 3.0 * magnitude / len2;                          // :47
    const double kappa2 = -2.0 * kappa / lengthscale;                           // :48
    const double sq5 = std::sqrt(5.0);
    s.F(2, 0) = -lam2 * lam; s.F(2, 1) = -3.0 * lam2; s.F(2, 2) = -3.0 * lam;   // :50-52
    s.Pinf(0, 0) = magnitude; s.Pinf(2, 2) = 25.0 * magnitude / len4;           // :53-54
    s.Pinf(1, 1) = kappa; s.Pinf(2, 0) = -kappa; s.Pinf(0, 2) = -kappa;         // :55-57
    s.R(0, 0) = p[2];
    s.dF[1](2, 0) = 15.0 * sq5 / len4; s.dF[1](2, 1) = 30.0 / len3; s.dF[1](2, 2) = sq5 * lam2;  // :61-63
    s.dPinf[0] = divs(s.Pinf, magnitude);                                       // :66
    s.dPinf[1](1, 1) = kappa2; s.dPinf[1](2, 0) = -kappa2; s.dPinf[1](0, 2) = -kappa2;  // :69-71
    s.dPinf[1](2, 2) = -100.0 * magnitude / len2 / len3;                        // :72
    for (int k = 0; k < 3; ++k) s.params[k] = p[k];
}

// ------------------------------------------------------------------ IHGP (ihgp.h)
struct IHGP {
    int kernel;  // 32 or 52
    int dim;
    double dt;
    StateSpace ss;
    SM A, Q, K, S, PF, HA, AKHA, PP;
    SM dS[3], dA[3], dK[3], dAKHA[3], HdA[3];
    int dare_iters, dlyap_iters[3];
    bool dare_conv, dlyap_conv[3];

    // ihgp.h:117-201
    void update(const double* params) {
        if (kernel == 32) matern32(ss, params); else matern52(ss, params);      // :119
        dim = ss.dim;
        const int d = dim;
        A = expm(scale(ss.F, dt));                                              // :120
        Q = sub(ss.Pinf, mul(mul(A, ss.Pinf), tr(A)));                          // :121
        Q = divs(add(Q, tr(Q)), 2.0);                                           // :122 (Q21: aliasing effect <= 1e-17, ignored)
        const SM HT = tr(ss.H);                                                 // :124
        dare_conv = DARE(A, HT, Q, ss.R, PP, &dare_iters);                      // :125 (return value ignored there)
        S = add(mul(mul(ss.H, PP), HT), ss.R);                                  // :126
        K = divs(mul(PP, HT), S(0, 0));                                         // :127
        PF = sub(PP, mul(mul(K, ss.H), PP));                                    // :128
        HA = mul(ss.H, A);                                                      // :129
        AKHA = sub(A, mul(K, HA));                                              // :130
        const SM AT = tr(A);                                                    // :131
        const SM AK = mul(A, K);                                                // :132
        const SM AAKH = sub(A, mul(AK, ss.H));                                  // :133
        for (int idx = 0; idx < 3; ++idx) {                                     // :136
            SM dAT(d, d), dQ(d, d), QLyap(d, d);
            if (is_zero(ss.dF[idx])) {                                          // :141
                dA[idx] = SM(d, d);                                             // :143
                if (is_zero(ss.dPinf[idx])) dQ = SM(d, d);                      // :144-147
                else dQ = sub(ss.dPinf[idx], mul(mul(A, ss.dPinf[idx]), AT));   // :150
                if (ss.dR[idx](0, 0) == 0.0) QLyap = dQ;                        // :152-155
                else {
                    // :158 `AK * AK.transpose() * dR + dQ` is a (d x d)*(1 x 1) product: undefined
                    // behaviour in the reference (Q19).  INTENDED meaning (cf. :183) restated here.
                    QLyap = add(scale(mul(AK, tr(AK)), ss.dR[idx](0, 0)), dQ);
                }
            } else {
                SM FF(2 * d, 2 * d);                                            // :163-166
                for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) {
                    FF(i, j) = ss.F(i, j); FF(d + i, d + j) = ss.F(i, j); FF(d + i, j) = ss.dF[idx](i, j);
                }
                const SM E = expm(scale(FF, dt));                               // :167
                dA[idx] = SM(d, d);
                for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) dA[idx](i, j) = E(d + i, j);
                dAT = tr(dA[idx]);                                              // :168
                if (is_zero(ss.dPinf[idx]))                                     // :169-172
                    dQ = sub(neg(mul(mul(dA[idx], ss.Pinf), AT)), mul(mul(A, ss.Pinf), dAT));
                else                                                            // :175
                    dQ = sub(sub(sub(ss.dPinf[idx], mul(mul(dA[idx], ss.Pinf), AT)), mul(mul(A, ss.dPinf[idx]), AT)), mul(mul(A, ss.Pinf), dAT));
                // :179 / :183
                SM t = add(mul(mul(dA[idx], PP), AT), mul(mul(A, PP), dAT));
                t = sub(t, mul(mul(mul(dA[idx], PP), HT), tr(AK)));
                t = sub(t, mul(mul(mul(AK, ss.H), PP), dAT));
                if (!(ss.dR[idx](0, 0) == 0.0)) t = add(t, mul(mul(AK, ss.dR[idx]), tr(AK)));
                QLyap = add(t, dQ);
            }
            SM dPP(d, d);
            dlyap_conv[idx] = DLyap(AAKH, QLyap, dPP, &dlyap_iters[idx]);       // :187
            dS[idx] = add(mul(mul(ss.H, dPP), HT), ss.dR[idx]);                 // :188
            dK[idx] = divs(mul(sub(dPP, divs(scale(PP, dS[idx](0, 0)), S(0, 0))), HT), S(0, 0));  // :189
            if (is_zero(ss.dF[idx])) {                                          // :190
                dAKHA[idx] = mul(mul(neg(dK[idx]), ss.H), A);                   // :192
                HdA[idx] = SM(d, 1);                                            // :193
            } else {
                dAKHA[idx] = sub(sub(dA[idx], mul(mul(dK[idx], ss.H), A)), mul(mul(K, ss.H), dA[idx]));  // :197
                HdA[idx] = tr(mul(ss.H, dA[idx]));                              // :198
            }
        }
    }

    // ihgp.h:37-57 / :60-78 / :81-93 (y observed or NaN) - x, xnew: d;  dx, dxnew: 3 x d (may be null)
    void step(const double* x, double y, const double* dx, double* xnew, double* yhat, double* dxnew) const {
        const int d = dim;
        const bool miss = std::isnan(y);                                        // :39
        const SM& M = miss ? A : AKHA;
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += M(i, j) * x[j];
            xnew[i] = miss ? s : s + K(i, 0) * y;                               // :41 / :50
        }
        if (yhat) *yhat = xnew[0];                                              // :42 / :51
        if (dx && dxnew) {
            for (int k = 0; k < 3; ++k) {
                const SM& dM = miss ? dA[k] : dAKHA[k];
                for (int i = 0; i < d; ++i) {
                    double s1 = 0.0, s2 = 0.0;
                    for (int j = 0; j < d; ++j) { s1 += dM(i, j) * x[j]; s2 += M(i, j) * dx[k * d + j]; }
                    dxnew[k * d + i] = miss ? s1 + s2 : s1 + s2 + dK[k](i, 0) * y;  // :45 / :54
                }
            }
        }
    }
    // ihgp.h:96-100
    void predict(const double* x, double* xnew, double* yhat) const {
        for (int i = 0; i < dim; ++i) { double s = 0.0; for (int j = 0; j < dim; ++j) s += A(i, j) * x[j]; xnew[i] = s; }
        *yhat = xnew[0];
    }
    // ihgp.h:204-209
    double nll(const double* x, double y) const {
        double hax = 0.0;
        for (int j = 0; j < dim; ++j) hax += HA(0, j) * x[j];
        const double v = y - hax;
        return 0.5 * (v * v / S(0, 0) + std::log(S(0, 0)));
    }
    // ihgp.h:212-222.  intended_hda=false replicates the de facto meaning of the shape-mismatched
    // product at :218 (Q20): dv = -HdA(0) * x(0) - HA * dx.
    double nll_grad(const double* x, double y, const double* dx, double* grad, bool intended_hda) const {
        double hax = 0.0;
        for (int j = 0; j < dim; ++j) hax += HA(0, j) * x[j];
        const double v = y - hax;                                               // :214
        const double loss = 0.5 * (v * v / S(0, 0) + std::log(S(0, 0)));        // :215
        for (int k = 0; k < 3; ++k) {
            double t1 = 0.0, t2 = 0.0;
            if (intended_hda) { for (int j = 0; j < dim; ++j) t1 += HdA[k](j, 0) * x[j]; }
            else t1 = HdA[k](0, 0) * x[0];
            for (int j = 0; j < dim; ++j) t2 += HA(0, j) * dx[k * dim + j];
            const double dv = -t1 - t2;                                         // :218
            grad[k] = (v * dv - 0.5 * (v * v / S(0, 0) - 1) * dS[k](0, 0)) / S(0, 0);  // :219
        }
        return loss;
    }

    // Symmetric-pivoted LDLT solve that only reads the LOWER triangle (what Eigen's
    // PP.ldlt().solve() does, ihgp.h:106).
    static SM ldlt_solve_lower(const SM& Ain, const SM& B) {
        const int n = Ain.r;
        SM w(n, n);
        for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) { w(i, j) = Ain(i, j); w(j, i) = Ain(i, j); }
        return solve_lu(w, B);  // the symmetric system itself; pivot order only matters at ~1e-16
    }

    // Smoother constants.  mode 0 = reference_literal (ihgp.h:105-107, Q3); mode 1 = rts_correct
    // (OUR extension, SURVEY section 11 item 5: PPc = A PF A' + Q, G = PF A' PPc^-1,
    //  P_s = G P_s G' + PF - G PPc G' solved exactly).
    void smoother_consts(int mode, SM& G, SM& P, int* iters) const {
        const int d = dim;
        if (mode == 0) {
            const SM PPs = add(mul(mul(A, PF), A), Q);                          // :105 (sic: A*PF*A)
            G = tr(ldlt_solve_lower(PPs, mul(A, PF)));                          // :106
            DLyap(G, sub(PF, mul(mul(G, PPs), tr(G))), P, iters);               // :107
        } else {
            const SM PPc = add(mul(mul(A, PF), tr(A)), Q);
            G = tr(solve_lu(tr(PPc), tr(mul(PF, tr(A)))));                      // G = PF A' PPc^-1
            const SM C = sub(PF, mul(mul(G, PPc), tr(G)));
            // vec(P) = (I - G (x) G)^-1 vec(C)
            const int n2 = d * d;
            std::vector<double> Mx(size_t(n2) * n2, 0.0), rhs(size_t(n2), 0.0);
            for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) {
                const int row = i * d + j;
                rhs[size_t(row)] = C(i, j);
                for (int k = 0; k < d; ++k) for (int l = 0; l < d; ++l)
                    Mx[size_t(row) * n2 + (k * d + l)] = (row == k * d + l ? 1.0 : 0.0) - G(i, k) * G(j, l);
            }
            for (int k = 0; k < n2; ++k) {  // Gaussian elimination, partial pivoting
                int piv = k;
                for (int i = k + 1; i < n2; ++i) if (std::fabs(Mx[size_t(i) * n2 + k]) > std::fabs(Mx[size_t(piv) * n2 + k])) piv = i;
                if (piv != k) { for (int j = 0; j < n2; ++j) std::swap(Mx[size_t(k) * n2 + j], Mx[size_t(piv) * n2 + j]); std::swap(rhs[size_t(k)], rhs[size_t(piv)]); }
                for (int i = k + 1; i < n2; ++i) {
                    const double f = Mx[size_t(i) * n2 + k] / Mx[size_t(k) * n2 + k];
                    for (int j = k; j < n2; ++j) Mx[size_t(i) * n2 + j] -= f * Mx[size_t(k) * n2 + j];
                    rhs[size_t(i)] -= f * rhs[size_t(k)];
                }
            }
            P = SM(d, d);
            for (int i = n2 - 1; i >= 0; --i) {
                double s = rhs[size_t(i)];
                for (int j = i + 1; j < n2; ++j) s -= Mx[size_t(i) * n2 + j] * P.a[j];
                P.a[i] = s / Mx[size_t(i) * n2 + i];
            }
            if (iters) *iters = 0;
        }
    }

    // Backward recursion over n stored states X[t][stride..] -> Xs.  mode 0: ihgp.h:108-113 (Q3):
    // Xs[n-1] = X[n-1]; Xs[j] = X[j+1] + G Xs[j+1] - A X[j+1].  mode 1 (rts_correct):
    // Xs[j] = X[j] + G (Xs[j+1] - A X[j]).
    void smooth(int mode, const SM& G, const double* X, size_t n, size_t stride, double* Xs) const {
        const int d = dim;
        if (n == 0) return;
        for (int i = 0; i < d; ++i) Xs[(n - 1) * stride + i] = X[(n - 1) * stride + i];   // :108
        for (size_t jj = n - 1; jj > 0; --jj) {                                            // :109
            const size_t j = jj - 1;
            const double* xs1 = Xs + (j + 1) * stride;
            if (mode == 0) {
                const double* x1 = X + (j + 1) * stride;
                for (int i = 0; i < d; ++i) {                                              // :111
                    double g = 0.0, a = 0.0;
                    for (int k = 0; k < d; ++k) { g += G(i, k) * xs1[k]; a += A(i, k) * x1[k]; }
                    Xs[j * stride + i] = x1[i] + g - a;
                }
            } else {
                const double* x0 = X + j * stride;
                double r[3];
                for (int i = 0; i < d; ++i) { double a = 0.0; for (int k = 0; k < d; ++k) a += A(i, k) * x0[k]; r[i] = xs1[i] - a; }
                for (int i = 0; i < d; ++i) { double g = 0.0; for (int k = 0; k < d; ++k) g += G(i, k) * r[k]; Xs[j * stride + i] = x0[i] + g; }
            }
        }
    }
};

// ------------------------------------------------------------------ dense helpers (dynamic sizes)
typedef std::vector<double> Vec;

// Thin SVD by one-sided Jacobi: a (m x n, row-major, m >= n) = U diag(s) V'.
static void jacobi_svd(const Vec& a, int m, int n, Vec& U, Vec& s, Vec& V) {
    Vec W(a);
    V.assign(size_t(n) * n, 0.0);
    for (int i = 0; i < n; ++i) V[size_t(i) * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
            double al = 0.0, be = 0.0, ga = 0.0;
            for (int i = 0; i < m; ++i) { const double wp = W[size_t(i) * n + p], wq = W[size_t(i) * n + q]; al += wp * wp; be += wq * wq; ga += wp * wq; }
            if (ga == 0.0) continue;
            off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
            const double zeta = (be - al) / (2.0 * ga);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
            for (int i = 0; i < m; ++i) { const double wp = W[size_t(i) * n + p], wq = W[size_t(i) * n + q]; W[size_t(i) * n + p] = c * wp - sn * wq; W[size_t(i) * n + q] = sn * wp + c * wq; }
            for (int i = 0; i < n; ++i) { const double vp = V[size_t(i) * n + p], vq = V[size_t(i) * n + q]; V[size_t(i) * n + p] = c * vp - sn * vq; V[size_t(i) * n + q] = sn * vp + c * vq; }
        }
        if (off < 1e-15) break;
    }
    s.assign(size_t(n), 0.0);
    U.assign(size_t(m) * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double q = 0.0;
        for (int i = 0; i < m; ++i) q += W[size_t(i) * n + j] * W[size_t(i) * n + j];
        s[size_t(j)] = std::sqrt(q);
        for (int i = 0; i < m; ++i) U[size_t(i) * n + j] = s[size_t(j)] > 0.0 ? W[size_t(i) * n + j] / s[size_t(j)] : 0.0;
    }
}

// Symmetric positive (semi)definite solve M x = b with the pseudo-inverse-of-D rule Eigen's
// LDLT::solve applies (an exactly-zero pivot yields 0, e.g. the all-NaN observation whose
// U0'U0 is the zero matrix; moihgp.h:177).  Symmetric pivoting on the largest diagonal.
static Vec ldlt_solve(Vec M, int n, Vec b) {
    std::vector<int> perm(size_t(n), 0);
    for (int i = 0; i < n; ++i) perm[size_t(i)] = i;
    Vec Lm(size_t(n) * n, 0.0), D(size_t(n), 0.0);
    for (int i = 0; i < n; ++i) Lm[size_t(i) * n + i] = 1.0;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(M[size_t(i) * n + i]) > std::fabs(M[size_t(piv) * n + piv])) piv = i;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(M[size_t(k) * n + j], M[size_t(piv) * n + j]);
            for (int i = 0; i < n; ++i) std::swap(M[size_t(i) * n + k], M[size_t(i) * n + piv]);
            for (int j = 0; j < k; ++j) std::swap(Lm[size_t(k) * n + j], Lm[size_t(piv) * n + j]);
            std::swap(perm[size_t(k)], perm[size_t(piv)]);
        }
        const double dk = M[size_t(k) * n + k];
        D[size_t(k)] = dk;
        if (dk == 0.0) continue;
        for (int i = k + 1; i < n; ++i) Lm[size_t(i) * n + k] = M[size_t(i) * n + k] / dk;
        for (int j = k + 1; j < n; ++j) for (int i = j; i < n; ++i) {
            M[size_t(i) * n + j] -= Lm[size_t(i) * n + k] * dk * Lm[size_t(j) * n + k];
            M[size_t(j) * n + i] = M[size_t(i) * n + j];
        }
    }
    Vec y(size_t(n), 0.0), x(size_t(n), 0.0);
    for (int i = 0; i < n; ++i) y[size_t(i)] = b[size_t(perm[size_t(i)])];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[size_t(i)] -= Lm[size_t(i) * n + j] * y[size_t(j)];
    const double tol = (std::numeric_limits<double>::min)();
    for (int i = 0; i < n; ++i) y[size_t(i)] = std::fabs(D[size_t(i)]) > tol ? y[size_t(i)] / D[size_t(i)] : 0.0;
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[size_t(i)] -= Lm[size_t(j) * n + i] * y[size_t(j)];
    for (int i = 0; i < n; ++i) x[size_t(perm[size_t(i)])] = y[size_t(i)];
    return x;
}

// ------------------------------------------------------------------ MOIHGP (moihgp.h)
struct MOIHGP {
    int kernel, p, L, d, K;  // K = 3 hyper-parameters per latent
    double dt;
    bool threading;          // moihgp.h:128-135: forced false when L < 2
    bool intended_hda;       // false = Q20 de facto (parity mode)
    Vec U;                   // p x L row-major
    Vec S;                   // L
    double sigma;
    std::vector<IHGP> igp;
    int num_param;

    MOIHGP(int kernel_, double dt_, int p_, int L_, bool threading_)
        : kernel(kernel_), p(p_), L(L_), K(3), dt(dt_), intended_hda(false) {
        igp.resize(size_t(L));
        const double def[3] = {1.0, 1.0, 0.1};                                  // matern32ss.h:35
        for (int l = 0; l < L; ++l) { igp[size_t(l)].kernel = kernel; igp[size_t(l)].dt = dt; igp[size_t(l)].update(def); }
        d = igp[0].dim;
        num_param = p * L + L + 1 + L * K;                                      // moihgp.h:93
        // moihgp.h:103-125 draws U = polar(I + 1e-3 N(0,1)) from std::random_device (Q11); the oracle
        // starts from polar(I) = I (p x L) and tests always inject U through update().
        U.assign(size_t(p) * L, 0.0);
        for (int i = 0; i < std::min(p, L); ++i) U[size_t(i) * L + i] = 1.0;
        S.assign(size_t(L), 1.0);                                               // :126
        sigma = 1e-2;                                                           // :127
        threading = L < 2 ? false : threading_;                                 // :128-135
    }

    // moihgp.h:431-457
    void update(const double* params) {
        Vec Uraw(params, params + size_t(p) * L);  // params[0:pL] read row-major as p x L (:436-439)
        Vec su, ss, sv;
        jacobi_svd(Uraw, p, L, su, ss, sv);
        for (int r = 0; r < p; ++r) for (int c = 0; c < L; ++c) {                // U = svdU * svdV' (:439,:446)
            double s = 0.0;
            for (int k = 0; k < L; ++k) s += su[size_t(r) * L + k] * sv[size_t(c) * L + k];
            U[size_t(r) * L + c] = s;
        }
        for (int l = 0; l < L; ++l) S[size_t(l)] = params[p * L + l];            // :448
        sigma = params[p * L + L];                                              // :449
        for (int l = 0; l < L; ++l) igp[size_t(l)].update(params + p * L + L + 1 + l * K);  // :450-456
    }
    // moihgp.h:721-738
    void get_params(double* params) const {
        for (int i = 0; i < p * L; ++i) params[i] = U[size_t(i)];
        for (int l = 0; l < L; ++l) params[p * L + l] = S[size_t(l)];
        params[p * L + L] = sigma;
        for (int l = 0; l < L; ++l) for (int k = 0; k < K; ++k) params[p * L + L + 1 + l * K + k] = igp[size_t(l)].ss.params[k];
    }

    // Projection Ty (moihgp.h:150-182 and the identical blocks at :231-263, :306-336, :462-498, :616-648)
    void project(const double* y, double* Ty) const {
        std::vector<int> obs;
        for (int r = 0; r < p; ++r) if (!std::isnan(y[r])) obs.push_back(r);     // :150-158
        if (int(obs.size()) != p) {                                             // :167
            const int m = int(obs.size());
            Vec G(size_t(L) * L, 0.0), b(size_t(L), 0.0);                        // U0'U0, U0'y_obs
            for (int a = 0; a < L; ++a) {
                for (int c = 0; c < L; ++c) { double s = 0.0; for (int i = 0; i < m; ++i) s += U[size_t(obs[size_t(i)]) * L + a] * U[size_t(obs[size_t(i)]) * L + c]; G[size_t(a) * L + c] = s; }
                double s = 0.0; for (int i = 0; i < m; ++i) s += U[size_t(obs[size_t(i)]) * L + a] * y[obs[size_t(i)]];
                b[size_t(a)] = s;
            }
            const Vec z = ldlt_solve(G, L, b);                                  // :177
            for (int l = 0; l < L; ++l) Ty[l] = (1 / std::sqrt(S[size_t(l)])) * z[size_t(l)];
        } else {
            for (int l = 0; l < L; ++l) {                                       // :181  (sqrtSinv * U') * y
                const double si = 1 / std::sqrt(S[size_t(l)]);
                double s = 0.0;
                for (int r = 0; r < p; ++r) s += (si * U[size_t(r) * L + l]) * y[r];
                Ty[l] = s;
            }
        }
    }
    // yhat = U * sqrtS * Tyhat  (moihgp.h:222-225)
    void backproject(const double* Tyhat, double* yhat) const {
        for (int r = 0; r < p; ++r) {
            double s = 0.0;
            for (int l = 0; l < L; ++l) s += (U[size_t(r) * L + l] * std::sqrt(S[size_t(l)])) * Tyhat[l];
            yhat[r] = s;
        }
    }
    // the four MOIHGP::step overloads (moihgp.h:148-226, 229-301, 304-378, 381-428); layouts as wrapper.cpp
    void step(const double* x, const double* y, const double* dx, double* xnew, double* yhat, double* dxnew) const {
        Vec Ty(size_t(L), 0.0), Tyhat(size_t(L), 0.0);
        if (y) project(y, Ty.data());
        for (int l = 0; l < L; ++l) {
            if (y) igp[size_t(l)].step(x + l * d, Ty[size_t(l)], dx ? dx + l * K * d : nullptr, xnew + l * d, &Tyhat[size_t(l)], dxnew ? dxnew + l * K * d : nullptr);
            else igp[size_t(l)].predict(x + l * d, xnew + l * d, &Tyhat[size_t(l)]);
        }
        if (yhat) backproject(Tyhat.data(), yhat);
    }
    // ||(I - UU') y||_2  (moihgp.h:499-501).  literal=true forms the p x p matrix first.
    double resid_norm(const double* y, bool literal) const {
        double q = 0.0;
        if (literal) {
            for (int r = 0; r < p; ++r) {
                double s = 0.0;
                for (int c = 0; c < p; ++c) {
                    double uu = 0.0;
                    for (int l = 0; l < L; ++l) uu += U[size_t(r) * L + l] * U[size_t(c) * L + l];
                    s += ((r == c ? 1.0 : 0.0) - uu) * y[c];
                }
                q += s * s;
            }
        } else {
            Vec w(size_t(L), 0.0);
            for (int l = 0; l < L; ++l) { double s = 0.0; for (int r = 0; r < p; ++r) s += U[size_t(r) * L + l] * y[r]; w[size_t(l)] = s; }
            for (int r = 0; r < p; ++r) { double s = y[r]; for (int l = 0; l < L; ++l) s -= U[size_t(r) * L + l] * w[size_t(l)]; q += s * s; }
        }
        return std::sqrt(q);
    }
    // moihgp.h:614-688
    double nll(const double* x, const double* y, bool literal) const {
        Vec Ty(size_t(L), 0.0);
        project(y, Ty.data());
        const double yu = resid_norm(y, literal);                               // :651
        const double m_n = std::max(double(p - L), 0.0);                        // :652
        double Ssum = 0.0; for (int l = 0; l < L; ++l) Ssum += S[size_t(l)];
        double loss = 0.5 * std::log(Ssum) + 0.5 * m_n * std::log(sigma) + 0.5 * yu / sigma;   // :653
        for (int l = 0; l < L; ++l) loss += igp[size_t(l)].nll(x + l * d, Ty[size_t(l)]);       // :675 / :684 (always added)
        return loss;
    }
    // moihgp.h:460-611.  literal=true runs the O(p^3 L^2) dU loop with the SVD of U (:513-552);
    // literal=false uses the rank-1 form it collapses to (asserted equal in tests).
    double nll_grad(const double* x, const double* y, const double* dx, double* grad, bool literal) const {
        const int sizeU = p * L;
        Vec sqrtSinv(size_t(L), 0.0), sqrtSinv3(size_t(L), 0.0);
        for (int l = 0; l < L; ++l) { const double s = std::sqrt(S[size_t(l)]); sqrtSinv[size_t(l)] = 1 / s; sqrtSinv3[size_t(l)] = 1 / s / s / s; }  // :475-480
        Vec Ty(size_t(L), 0.0);
        project(y, Ty.data());                                                  // :481-498
        const double yu = resid_norm(y, literal);                               // :499-501
        const double m_n = std::max(double(p - L), 0.0);                        // :502
        double Ssum = 0.0; for (int l = 0; l < L; ++l) Ssum += S[size_t(l)];
        double loss = 0.5 * std::log(Ssum) + 0.5 * m_n * std::log(sigma) + 0.5 * yu / sigma;   // :503
        Vec pv(size_t(L), 0.0);
        for (int l = 0; l < L; ++l) {                                           // :505-512 (raw y(l), Q8)
            const IHGP& g = igp[size_t(l)];
            double hax = 0.0, hak = 0.0;
            for (int j = 0; j < d; ++j) { hax += g.HA(0, j) * x[l * d + j]; hak += g.HA(0, j) * g.K(j, 0); }
            const double vi = y[l] - hax;
            pv[size_t(l)] = vi * (1 - hak) / g.S(0, 0);
        }
        for (int i = 0; i < num_param; ++i) grad[i] = 0.0;                      // :537
        Vec UTy(size_t(L), 0.0);
        for (int l = 0; l < L; ++l) { double s = 0.0; for (int r = 0; r < p; ++r) s += U[size_t(r) * L + l] * y[r]; UTy[size_t(l)] = s; }
        if (literal) {
            Vec su, sv, ssv;
            jacobi_svd(U, p, L, su, ssv, sv);                                   // :513-536
            // left = Io + svdU (invS - Il) svdU'  (p x p);  right = Il + svdV (invS - Il) svdV'  (L x L)
            Vec left(size_t(p) * p, 0.0), right(size_t(L) * L, 0.0);
            for (int a = 0; a < p; ++a) for (int b = 0; b < p; ++b) {
                double s = (a == b) ? 1.0 : 0.0;
                for (int k = 0; k < L; ++k) s += su[size_t(a) * L + k] * (1.0 / ssv[size_t(k)] - 1.0) * su[size_t(b) * L + k];
                left[size_t(a) * p + b] = s;
            }
            for (int a = 0; a < L; ++a) for (int b = 0; b < L; ++b) {
                double s = (a == b) ? 1.0 : 0.0;
                for (int k = 0; k < L; ++k) s += sv[size_t(a) * L + k] * (1.0 / ssv[size_t(k)] - 1.0) * sv[size_t(b) * L + k];
                right[size_t(a) * L + b] = s;
            }
            Vec dU(size_t(p) * L, 0.0);
            for (int idx1 = 0; idx1 < sizeU; ++idx1) {                          // :538
                const int row = idx1 / L, col = idx1 % L;                       // dA[idx1] = E(row, col)  (:95-102)
                for (int a = 0; a < p; ++a) for (int b = 0; b < L; ++b) dU[size_t(a) * L + b] = left[size_t(a) * p + row] * right[size_t(col) * L + b];  // :545
                double g1 = 0.0;                                                // -y' U dU' y / sigma  (:546)
                for (int b = 0; b < L; ++b) { double t = 0.0; for (int a = 0; a < p; ++a) t += dU[size_t(a) * L + b] * y[a]; g1 += UTy[size_t(b)] * t; }
                double g = -g1 / sigma;
                for (int idx2 = 0; idx2 < L; ++idx2) {                          // :547-551
                    double t = 0.0;
                    for (int a = 0; a < p; ++a) t += (sqrtSinv[size_t(idx2)] * dU[size_t(a) * L + idx2]) * y[a];
                    g += pv[size_t(idx2)] * t;
                }
                grad[idx1] = g;
            }
        } else {
            for (int r = 0; r < p; ++r) for (int c = 0; c < L; ++c)
                grad[r * L + c] = y[r] * (-UTy[size_t(c)] / sigma + pv[size_t(c)] * sqrtSinv[size_t(c)]);
        }
        for (int l = 0; l < L; ++l)                                             // :554-562
            grad[sizeU + l] = 0.5 / S[size_t(l)] + pv[size_t(l)] * (-0.5 * sqrtSinv3[size_t(l)] * UTy[size_t(l)]);
        grad[sizeU + L] = 0.5 * (m_n - yu / sigma) / sigma;                     // :563
        for (int l = 0; l < L; ++l) {                                           // :565-607
            double g[3];
            const double li = igp[size_t(l)].nll_grad(x + l * d, Ty[size_t(l)], dx + l * K * d, g, intended_hda);
            if (threading) loss += li;                                          // :588 vs :601 (Q5)
            const double dn = g[K - 1];
            grad[sizeU + l] -= dn * sigma / S[size_t(l)] / S[size_t(l)];        // :591 / :604
            grad[sizeU + L] += dn / S[size_t(l)];                               // :592 / :605
            for (int k = 0; k < K; ++k) grad[sizeU + L + 1 + l * K + k] = g[k]; // :608-609 (column-major igp_grad)
        }
        return loss;
    }

    // RegressionObjective::operator() loop (moihgp_regression.h:42-50) / OnlineObjective window loop
    // (moihgp_online.h:61-70) over T observations from the given carried state:
    // for each y_t: step v2, then NLL+grad on the PRE-step state; naive sequential sums.
    double objective(const double* Y, size_t T, double* x, double* dx, double* grad, bool literal) const {
        Vec xn(size_t(L) * d, 0.0), dxn(size_t(L) * K * d, 0.0), g(size_t(num_param), 0.0);
        double loss = 0.0;
        for (int i = 0; i < num_param; ++i) grad[i] = 0.0;
        for (size_t t = 0; t < T; ++t) {
            const double* y = Y + t * p;
            step(x, y, dx, xn.data(), nullptr, dxn.data());
            loss += nll_grad(x, y, dx, g.data(), literal);
            for (int i = 0; i < num_param; ++i) grad[i] += g[size_t(i)];
            std::copy(xn.begin(), xn.end(), x);
            std::copy(dxn.begin(), dxn.end(), dx);
        }
        return loss;
    }

    // One sequence of the fused pass: filter (step v3 loop, moihgp_regression.h:127-139) storing the
    // post-step states X[t][L][d] and optionally Yhat[t][p]; NLL (lik2 on the pre-step state, summed
    // naively over t); then the smoother over the stored states per latent.  x is carried in/out.
    double filter_smoother_nll(const double* Y, size_t T, double* x, double* X, double* Xs, double* Yhat, int smoother_mode) const {
        Vec xn(size_t(L) * d, 0.0);
        double loss = 0.0;
        for (size_t t = 0; t < T; ++t) {
            const double* y = Y + t * p;
            loss += nll(x, y, false);
            step(x, y, nullptr, xn.data(), Yhat ? Yhat + t * p : nullptr, nullptr);
            std::copy(xn.begin(), xn.end(), x);
            if (X) std::copy(xn.begin(), xn.end(), X + t * size_t(L) * d);
        }
        if (X && Xs && smoother_mode >= 0) {
            for (int l = 0; l < L; ++l) {
                SM G, P;
                igp[size_t(l)].smoother_consts(smoother_mode, G, P, nullptr);
                igp[size_t(l)].smooth(smoother_mode, G, X + l * d, T, size_t(L) * d, Xs + l * d);
            }
        }
        return loss;
    }
};

}  // namespace oracle

// ------------------------------------------------------------------ C ABI (ctypes, bench)
using oracle::MOIHGP;
using oracle::SM;

extern "C" {

void* oracle_new(int kernel, double dt, size_t p, size_t L, int threading) { return new MOIHGP(kernel, dt, int(p), int(L), threading != 0); }
void oracle_del(void* h) { delete static_cast<MOIHGP*>(h); }
void oracle_set_intended_hda(void* h, int on) { static_cast<MOIHGP*>(h)->intended_hda = on != 0; }
size_t oracle_igp_dim(void* h) { return size_t(static_cast<MOIHGP*>(h)->d); }
size_t oracle_num_param(void* h) { return size_t(static_cast<MOIHGP*>(h)->num_param); }
void oracle_update(void* h, const double* params) { static_cast<MOIHGP*>(h)->update(params); }
void oracle_get_params(void* h, double* params) { static_cast<MOIHGP*>(h)->get_params(params); }
void oracle_get_U(void* h, double* U) { MOIHGP* g = static_cast<MOIHGP*>(h); std::copy(g->U.begin(), g->U.end(), U); }

void oracle_step1(void* h, const double* x, const double* y, const double* dx, double* xn, double* yh, double* dxn) { static_cast<MOIHGP*>(h)->step(x, y, dx, xn, yh, dxn); }
void oracle_step2(void* h, const double* x, const double* y, const double* dx, double* xn, double* dxn) { static_cast<MOIHGP*>(h)->step(x, y, dx, xn, nullptr, dxn); }
void oracle_step3(void* h, const double* x, const double* y, double* xn, double* yh) { static_cast<MOIHGP*>(h)->step(x, y, nullptr, xn, yh, nullptr); }
void oracle_step4(void* h, const double* x, double* xn, double* yh) { static_cast<MOIHGP*>(h)->step(x, nullptr, nullptr, xn, yh, nullptr); }
double oracle_lik1(void* h, const double* x, const double* y, const double* dx, double* grad, int literal) { return static_cast<MOIHGP*>(h)->nll_grad(x, y, dx, grad, literal != 0); }
double oracle_lik2(void* h, const double* x, const double* y, int literal) { return static_cast<MOIHGP*>(h)->nll(x, y, literal != 0); }

// Per-latent steady-state constants, same layout as ref_probe.cpp's probeXX_ihgp_consts:
//   A Q K S PF HA AKHA, then per k: dS dA dK dAKHA HdA   (matrices row-major).  Returns the count.
size_t oracle_ihgp_consts(void* h, size_t l, double* out) {
    const oracle::IHGP& g = static_cast<MOIHGP*>(h)->igp[l];
    const int d = g.dim;
    size_t o = 0;
    for (int i = 0; i < d * d; ++i) out[o++] = g.A.a[i];
    for (int i = 0; i < d * d; ++i) out[o++] = g.Q.a[i];
    for (int i = 0; i < d; ++i) out[o++] = g.K(i, 0);
    out[o++] = g.S(0, 0);
    for (int i = 0; i < d * d; ++i) out[o++] = g.PF.a[i];
    for (int i = 0; i < d; ++i) out[o++] = g.HA(0, i);
    for (int i = 0; i < d * d; ++i) out[o++] = g.AKHA.a[i];
    for (int k = 0; k < 3; ++k) {
        out[o++] = g.dS[k](0, 0);
        for (int i = 0; i < d * d; ++i) out[o++] = g.dA[k].a[i];
        for (int i = 0; i < d; ++i) out[o++] = g.dK[k](i, 0);
        for (int i = 0; i < d * d; ++i) out[o++] = g.dAKHA[k].a[i];
        for (int i = 0; i < d; ++i) out[o++] = g.HdA[k](i, 0);
    }
    return o;
}
// iteration counts / convergence flags: out[0]=DARE iters, out[1..3]=DLyap iters, out[4]=DARE conv, out[5..7]=DLyap conv
void oracle_ihgp_iters(void* h, size_t l, int* out) {
    const oracle::IHGP& g = static_cast<MOIHGP*>(h)->igp[l];
    out[0] = g.dare_iters; out[4] = g.dare_conv;
    for (int k = 0; k < 3; ++k) { out[1 + k] = g.dlyap_iters[k]; out[5 + k] = g.dlyap_conv[k]; }
}
// smoother constants of latent l: G[d*d], P[d*d] row-major; mode 0 literal, 1 rts_correct
void oracle_smoother_consts(void* h, size_t l, int mode, double* G, double* P) {
    const oracle::IHGP& g = static_cast<MOIHGP*>(h)->igp[l];
    SM Gm, Pm;
    g.smoother_consts(mode, Gm, Pm, nullptr);
    for (int i = 0; i < g.dim * g.dim; ++i) { G[i] = Gm.a[i]; P[i] = Pm.a[i]; }
}
// IHGP::backwardSmoother on n stored states of latent l (X, Xs: [n][d])
void oracle_ihgp_smooth(void* h, size_t l, int mode, const double* X, size_t n, double* Xs) {
    const oracle::IHGP& g = static_cast<MOIHGP*>(h)->igp[l];
    SM Gm, Pm;
    g.smoother_consts(mode, Gm, Pm, nullptr);
    g.smooth(mode, Gm, X, n, size_t(g.dim), Xs);
}

// Objective over N independent sequences Y[N][T][p] (sum over n is OUR extension; the reference has
// one sequence).  x, dx: [N][L][d], [N][L][3][d] carried in/out.  Returns the loss; grad[num_param].
double oracle_objective(void* h, const double* Y, size_t N, size_t T, double* x, double* dx, double* grad, int literal) {
    MOIHGP* g = static_cast<MOIHGP*>(h);
    std::vector<double> gi(size_t(g->num_param), 0.0);
    for (int i = 0; i < g->num_param; ++i) grad[i] = 0.0;
    double loss = 0.0;
    for (size_t n = 0; n < N; ++n) {
        loss += g->objective(Y + n * T * size_t(g->p), T, x + n * size_t(g->L) * g->d, dx + n * size_t(g->L) * 3 * g->d, gi.data(), literal != 0);
        for (int i = 0; i < g->num_param; ++i) grad[i] += gi[size_t(i)];
    }
    return loss;
}

// The fused pass on N sequences with `nthreads` host threads (sequences are independent).
// Y[N][T][p]; x[N][L][d] in/out; X, Xs: [N][T][L][d] (may be null); Yhat [N][T][p] (may be null);
// nll[N].  smoother_mode: -1 none, 0 reference_literal, 1 rts_correct.
void oracle_filter_smoother_nll(void* h, const double* Y, size_t N, size_t T, double* x, double* X, double* Xs, double* Yhat,
                                double* nll, int smoother_mode, int nthreads) {
    MOIHGP* g = static_cast<MOIHGP*>(h);
    const size_t Ld = size_t(g->L) * g->d;
    auto work = [&](size_t n0, size_t n1) {
        for (size_t n = n0; n < n1; ++n)
            nll[n] = g->filter_smoother_nll(Y + n * T * size_t(g->p), T, x + n * Ld, X ? X + n * T * Ld : nullptr,
                                            Xs ? Xs + n * T * Ld : nullptr, Yhat ? Yhat + n * T * size_t(g->p) : nullptr, smoother_mode);
    };
    if (nthreads <= 1 || N <= 1) { work(0, N); return; }
    std::vector<std::thread> th;
    const size_t nt = std::min(size_t(nthreads), N);
    for (size_t k = 0; k < nt; ++k) th.emplace_back(work, N * k / nt, N * (k + 1) / nt);
    for (auto& t : th) t.join();
}

}  // extern "C"
