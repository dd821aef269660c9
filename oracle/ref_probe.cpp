// TEST INFRASTRUCTURE ONLY (oracle/).  A harness of OUR OWN that #includes the UNMODIFIED
// reference headers where they lie (/root/reference/moihgp/include, compiled against
// oracle/eigen_shim) and exposes what the reference's C ABI (src/wrapper.cpp) does not:
//   * the Matern-5/2 instantiation MOIHGP<Matern52StateSpace>  (wrapper.cpp:22 typedefs GP52
//     to the Matern-3/2 class, SURVEY Q7, so gp52_* cannot reach it),
//   * IHGP's public steady-state members A,Q,K,S,PF,HA,AKHA,dS,dA,dK,dAKHA,HdA (ihgp.h:243-254),
//   * IHGP::backwardSmoother (ihgp.h:103-114), which has no caller in the reference.
// It is used only to pin the CPU restatement (moihgp_oracle.cpp) and to generate the golden
// fixtures under tests/golden/ (oracle/gen_golden.py), and - probeXX_run_pass, the -O2 flavour - as the CPU arm of
// bench.py (`--impl reference`, cpu_baseline kind "reference").  It is never shipped.
#include <cstddef>
#include <vector>
#include <Eigen/Core>
#include <moihgp/moihgp.h>
#include <moihgp/matern32ss.h>
#include <moihgp/matern52ss.h>

namespace {

typedef std::vector<Eigen::VectorXd> VecList;
typedef std::vector<std::vector<Eigen::VectorXd> > VecList2;

template <typename SS>
struct Probe {
    typedef moihgp::MOIHGP<SS> GP;

    static void load_x(GP* gp, const double* x, VecList& out) {
        const size_t L = gp->getNumLatent(), d = gp->getIGPDim();
        out.assign(L, Eigen::VectorXd(d));
        for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < d; ++i) out[l](i) = x[l * d + i];
    }
    static void load_dx(GP* gp, const double* dx, VecList2& out) {
        const size_t L = gp->getNumLatent(), d = gp->getIGPDim(), K = gp->getNumIGPParam();
        out.assign(L, VecList(K, Eigen::VectorXd(d)));
        for (size_t l = 0; l < L; ++l) for (size_t k = 0; k < K; ++k) for (size_t i = 0; i < d; ++i)
            out[l][k](i) = dx[(l * K + k) * d + i];
    }
    static void store_x(GP* gp, const VecList& in, double* x) {
        const size_t L = gp->getNumLatent(), d = gp->getIGPDim();
        for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < d; ++i) x[l * d + i] = in[l](i);
    }
    static void store_dx(GP* gp, const VecList2& in, double* dx) {
        const size_t L = gp->getNumLatent(), d = gp->getIGPDim(), K = gp->getNumIGPParam();
        for (size_t l = 0; l < L; ++l) for (size_t k = 0; k < K; ++k) for (size_t i = 0; i < d; ++i)
            dx[(l * K + k) * d + i] = in[l][k](i);
    }
    static Eigen::VectorXd load_y(GP* gp, const double* y) {
        Eigen::VectorXd v(gp->getNumOutput());
        for (size_t i = 0; i < gp->getNumOutput(); ++i) v(i) = y[i];
        return v;
    }

    static void step1(GP* gp, const double* x, const double* y, const double* dx, double* xnew, double* yhat, double* dxnew) {
        VecList X, Xn; VecList2 DX, DXn;
        load_x(gp, x, X); load_dx(gp, dx, DX); Xn = X; DXn = DX;
        Eigen::VectorXd Y = load_y(gp, y), Yh(gp->getNumOutput());
        gp->step(X, Y, DX, Xn, Yh, DXn);
        store_x(gp, Xn, xnew); store_dx(gp, DXn, dxnew);
        for (size_t i = 0; i < gp->getNumOutput(); ++i) yhat[i] = Yh(i);
    }
    static void step2(GP* gp, const double* x, const double* y, const double* dx, double* xnew, double* dxnew) {
        VecList X, Xn; VecList2 DX, DXn;
        load_x(gp, x, X); load_dx(gp, dx, DX); Xn = X; DXn = DX;
        Eigen::VectorXd Y = load_y(gp, y);
        gp->step(X, Y, DX, Xn, DXn);
        store_x(gp, Xn, xnew); store_dx(gp, DXn, dxnew);
    }
    static void step3(GP* gp, const double* x, const double* y, double* xnew, double* yhat) {
        VecList X, Xn;
        load_x(gp, x, X); Xn = X;
        Eigen::VectorXd Y = load_y(gp, y), Yh(gp->getNumOutput());
        gp->step(X, Y, Xn, Yh);
        store_x(gp, Xn, xnew);
        for (size_t i = 0; i < gp->getNumOutput(); ++i) yhat[i] = Yh(i);
    }
    static void step4(GP* gp, const double* x, double* xnew, double* yhat) {
        VecList X, Xn;
        load_x(gp, x, X); Xn = X;
        Eigen::VectorXd Yh(gp->getNumOutput());
        gp->step(X, Xn, Yh);
        store_x(gp, Xn, xnew);
        for (size_t i = 0; i < gp->getNumOutput(); ++i) yhat[i] = Yh(i);
    }
    static double lik1(GP* gp, const double* x, const double* y, const double* dx, double* grad) {
        VecList X; VecList2 DX;
        load_x(gp, x, X); load_dx(gp, dx, DX);
        Eigen::VectorXd Y = load_y(gp, y), G(gp->getNumParam());
        const double loss = gp->negLogLikelihood(X, Y, DX, G);
        for (size_t i = 0; i < gp->getNumParam(); ++i) grad[i] = G(i);
        return loss;
    }
    static double lik2(GP* gp, const double* x, const double* y) {
        VecList X;
        load_x(gp, x, X);
        Eigen::VectorXd Y = load_y(gp, y);
        return gp->negLogLikelihood(X, Y);
    }

    // IHGP steady-state members after update(params); `out` layout (d = dim, K = 3):
    //   A[d*d] Q[d*d] K[d] S[1] PF[d*d] HA[d] AKHA[d*d]  then per k: dS[1] dA[d*d] dK[d] dAKHA[d*d] HdA[d]
    // matrices row-major.
    static size_t ihgp_consts(double dt, const double* params, double* out) {
        moihgp::IHGP<SS> gp(dt);
        Eigen::VectorXd p(3);
        p(0) = params[0]; p(1) = params[1]; p(2) = params[2];
        gp.update(p);
        const size_t d = gp.getDim();
        size_t o = 0;
        for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.A(i, j);
        for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.Q(i, j);
        for (size_t i = 0; i < d; ++i) out[o++] = gp.K(i, 0);
        out[o++] = gp.S(0, 0);
        for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.PF(i, j);
        for (size_t i = 0; i < d; ++i) out[o++] = gp.HA(0, i);
        for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.AKHA(i, j);
        for (size_t k = 0; k < 3; ++k) {
            out[o++] = gp.dS[k](0, 0);
            for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.dA[k](i, j);
            for (size_t i = 0; i < d; ++i) out[o++] = gp.dK[k](i, 0);
            for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) out[o++] = gp.dAKHA[k](i, j);
            for (size_t i = 0; i < d; ++i) out[o++] = gp.HdA[k](i, 0);
        }
        return o;
    }

    // IHGP::backwardSmoother on n stored states X[n][d]; outputs Xs[n][d], P[d*d], G[d*d] row-major.
    static void ihgp_smoother(double dt, const double* params, const double* X, size_t n, double* Xs, double* P, double* G) {
        moihgp::IHGP<SS> gp(dt);
        Eigen::VectorXd p(3);
        p(0) = params[0]; p(1) = params[1]; p(2) = params[2];
        gp.update(p);
        const size_t d = gp.getDim();
        VecList Xin(n, Eigen::VectorXd(d)), Xout;
        for (size_t t = 0; t < n; ++t) for (size_t i = 0; i < d; ++i) Xin[t](i) = X[t * d + i];
        Eigen::MatrixXd Pm, Gm;
        gp.backwardSmoother(Xin, Xout, Pm, Gm);
        for (size_t t = 0; t < n; ++t) for (size_t i = 0; i < d; ++i) Xs[t * d + i] = Xout[t](i);
        for (size_t i = 0; i < d; ++i) for (size_t j = 0; j < d; ++j) { P[i * d + j] = Pm(i, j); G[i * d + j] = Gm(i, j); }
    }

    // The fused pass as the reference's own classes run it: the loop of MOIHGPRegression::predict
    // (moihgp_regression.h:127-139: step(x, y, xnew, yhat)) with negLogLikelihood(x, y) on the pre-step state
    // (moihgp.h:614-688), then IHGP::backwardSmoother (ihgp.h:103-114) per latent on the stored filtered states.
    // Y[T][p]; igp_params[L][3]; X, Xs [T][L][d] (may be null).  Returns the summed NLL.
    static double run_pass(GP* gp, double dt, const double* igp_params, const double* Y, size_t T, double* Xo, double* Xso, int smooth) {
        const size_t L = gp->getNumLatent(), d = gp->getIGPDim(), p = gp->getNumOutput();
        VecList X(L, Eigen::VectorXd(d)), Xn(L, Eigen::VectorXd(d));
        for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < d; ++i) X[l](i) = 0.0;
        Xn = X;
        std::vector<VecList> series(L, VecList(T, Eigen::VectorXd(d)));
        Eigen::VectorXd y(p), yh(p);
        double nll = 0.0;
        for (size_t t = 0; t < T; ++t) {
            for (size_t i = 0; i < p; ++i) y(i) = Y[t * p + i];
            nll += gp->negLogLikelihood(X, y);
            gp->step(X, y, Xn, yh);
            X = Xn;
            for (size_t l = 0; l < L; ++l) series[l][t] = X[l];
            if (Xo) for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < d; ++i) Xo[(t * L + l) * d + i] = X[l](i);
        }
        if (smooth) {
            for (size_t l = 0; l < L; ++l) {
                moihgp::IHGP<SS> ig(dt);
                Eigen::VectorXd q(3);
                q(0) = igp_params[3 * l]; q(1) = igp_params[3 * l + 1]; q(2) = igp_params[3 * l + 2];
                ig.update(q);
                VecList out;
                Eigen::MatrixXd Pm, Gm;
                ig.backwardSmoother(series[l], out, Pm, Gm);
                if (Xso) for (size_t t = 0; t < T; ++t) for (size_t i = 0; i < d; ++i) Xso[(t * L + l) * d + i] = out[t](i);
            }
        }
        return nll;
    }
};

typedef Probe<moihgp::Matern32StateSpace> P32;
typedef Probe<moihgp::Matern52StateSpace> P52;

}  // namespace

#define PROBE_API(XX, PXX)                                                                                      \
    void* probe##XX##_new(double dt, size_t p, size_t L, bool threading) { return new PXX::GP(dt, p, L, threading); } \
    void probe##XX##_del(void* gp) { delete static_cast<PXX::GP*>(gp); }                                        \
    void probe##XX##_update(void* gp, const double* params) {                                                   \
        PXX::GP* g = static_cast<PXX::GP*>(gp);                                                                 \
        Eigen::VectorXd v(g->getNumParam());                                                                    \
        for (size_t i = 0; i < g->getNumParam(); ++i) v(i) = params[i];                                         \
        g->update(v);                                                                                           \
    }                                                                                                           \
    void probe##XX##_get_params(void* gp, double* params) {                                                     \
        PXX::GP* g = static_cast<PXX::GP*>(gp);                                                                 \
        Eigen::VectorXd v = g->getParams();                                                                     \
        for (size_t i = 0; i < g->getNumParam(); ++i) params[i] = v(i);                                         \
    }                                                                                                           \
    void probe##XX##_get_U(void* gp, double* U) { /* row-major p x L */                                         \
        PXX::GP* g = static_cast<PXX::GP*>(gp);                                                                 \
        for (size_t r = 0; r < g->getNumOutput(); ++r) for (size_t c = 0; c < g->getNumLatent(); ++c)           \
            U[r * g->getNumLatent() + c] = g->U(r, c);                                                          \
    }                                                                                                           \
    size_t probe##XX##_igp_dim(void* gp) { return static_cast<PXX::GP*>(gp)->getIGPDim(); }                     \
    size_t probe##XX##_num_param(void* gp) { return static_cast<PXX::GP*>(gp)->getNumParam(); }                 \
    void probe##XX##_step1(void* gp, const double* x, const double* y, const double* dx, double* xn, double* yh, double* dxn) { PXX::step1(static_cast<PXX::GP*>(gp), x, y, dx, xn, yh, dxn); } \
    void probe##XX##_step2(void* gp, const double* x, const double* y, const double* dx, double* xn, double* dxn) { PXX::step2(static_cast<PXX::GP*>(gp), x, y, dx, xn, dxn); } \
    void probe##XX##_step3(void* gp, const double* x, const double* y, double* xn, double* yh) { PXX::step3(static_cast<PXX::GP*>(gp), x, y, xn, yh); } \
    void probe##XX##_step4(void* gp, const double* x, double* xn, double* yh) { PXX::step4(static_cast<PXX::GP*>(gp), x, xn, yh); } \
    double probe##XX##_lik1(void* gp, const double* x, const double* y, const double* dx, double* grad) { return PXX::lik1(static_cast<PXX::GP*>(gp), x, y, dx, grad); } \
    double probe##XX##_lik2(void* gp, const double* x, const double* y) { return PXX::lik2(static_cast<PXX::GP*>(gp), x, y); } \
    size_t probe##XX##_ihgp_consts(double dt, const double* params, double* out) { return PXX::ihgp_consts(dt, params, out); } \
    void probe##XX##_ihgp_smoother(double dt, const double* params, const double* X, size_t n, double* Xs, double* P, double* G) { PXX::ihgp_smoother(dt, params, X, n, Xs, P, G); } \
    double probe##XX##_run_pass(void* gp, double dt, const double* igp_params, const double* Y, size_t T, double* X, double* Xs, int smooth) { return PXX::run_pass(static_cast<PXX::GP*>(gp), dt, igp_params, Y, T, X, Xs, smooth); }

extern "C" {
PROBE_API(32, P32)
PROBE_API(52, P52)
}
