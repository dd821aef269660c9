// moihgp_b200/moihgp.hpp - host-side C++ mirror of the reference's class API for the hot path, over the C ABI of
// libmoihgp.so (include/moihgp_b200.h).  Header-only, no dependency beyond the C ABI.
//
// Same class and method names, argument meaning and (absent) error behaviour as the reference:
//   moihgp::MOIHGP<StateSpace>              moihgp/include/moihgp/moihgp.h:76-757
//   moihgp::RegressionObjective<StateSpace> moihgp/include/moihgp/moihgp_regression.h:17-70
//   moihgp::OnlineObjective<StateSpace>     moihgp/include/moihgp/moihgp_online.h:18-115
// living in namespace moihgp_b200 so that both header sets can be included side by side; a reference build switches
// with `namespace moihgp = moihgp_b200;`.
//
// The reference's vector type is Eigen::VectorXd.  Eigen is not a dependency here: every class takes the vector type
// as a template parameter `Vec` (default std::vector<double>) and only needs size(), resize(n), data() and operator[]
// - which Eigen::VectorXd provides - so `MOIHGP<Matern32StateSpace, Eigen::VectorXd>` gives the reference's exact
// signatures, and the functors plug into LBFGSpp::LBFGSBSolver::minimize(f, x, fx, lb, ub) unchanged.
//
// What runs where: the per-observation methods (step, negLogLikelihood) are one small kernel launch each (the legacy
// gpXX_* path); the objective functors hand the WHOLE window / data set to moihgp_cuda_objective - one fused device pass
// instead of the reference's loop of step + negLogLikelihood per observation.
#ifndef MOIHGP_B200_MOIHGP_HPP
#define MOIHGP_B200_MOIHGP_HPP

#include <cstddef>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include "../moihgp_b200.h"

namespace moihgp_b200 {

// StateSpace tags (the reference's duck-typed classes matern32ss.h:13-99 / matern52ss.h:13-110 reduce, on this side of
// the boundary, to the choice of kernel; their arithmetic runs in the K-setup kernel)
struct Matern32StateSpace { static constexpr int kernel = MOIHGP_MATERN32; static constexpr int dim = 2; };
struct Matern52StateSpace { static constexpr int kernel = MOIHGP_MATERN52; static constexpr int dim = 3; };

namespace detail {
template <typename Vec> inline Vec make_vec(size_t n) { Vec v; v.resize(n); for (size_t i = 0; i < n; ++i) v[i] = 0.0; return v; }
inline void check(moihgp_handle* h, int rc, const char* where) {
    if (rc != 0) throw std::runtime_error(std::string(where) + ": " + moihgp_cuda_last_error(h));
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------------------------
// MOIHGP<StateSpace>   (moihgp.h:76-757)
template <typename StateSpace, typename Vec = std::vector<double> >
class MOIHGP {
public:
    typedef std::vector<Vec> State;                      // x[l]      : IGP state of latent l            (d)
    typedef std::vector<std::vector<Vec> > DState;       // dx[l][k]  : its derivative w.r.t. parameter k (d)

    // moihgp.h:81-136.  U starts as a random near-identity polar factor like the reference's (moihgp.h:103-125) when
    // `random_U` (the default, as the reference), or as the exact identity block otherwise.
    MOIHGP(const double& dt, const size_t& num_output, const size_t& num_latent, const bool& threading, int device = -1,
           bool random_U = true)
        : _h(NULL), _num_output(num_output), _num_latent(num_latent) {
        if (random_U) {
            _h = StateSpace::kernel == MOIHGP_MATERN32 ? gp32_new(dt, num_output, num_latent, threading) : NULL;
        }
        if (!_h) {
            if (moihgp_cuda_create(&_h, StateSpace::kernel, dt, num_output, num_latent, threading ? 1 : 0, device) != 0 || !_h)
                throw std::runtime_error("moihgp_cuda_create failed (no B200-class CUDA device? there is no CPU fallback)");
            if (random_U) {
                // Matern-5/2 with the reference's random start: draw through a throw-away Matern-3/2 handle (same U recipe)
                moihgp_handle* tmp = gp32_new(dt, num_output, num_latent, threading);
                std::vector<double> p(moihgp_cuda_num_param(tmp));
                moihgp_cuda_get_params(tmp, p.data());
                moihgp_cuda_destroy(tmp);
                moihgp_cuda_update(_h, p.data());
            }
        }
        _dim = moihgp_cuda_igp_dim(_h);
        _num_param = moihgp_cuda_num_param(_h);
        _igp_num_param = moihgp_cuda_num_igp_param(_h);
        _xb.resize(2 * _num_latent * _dim);
        _dxb.resize(2 * _num_latent * _igp_num_param * _dim);
        _yb.resize(2 * _num_output);
        _pb.resize(_num_param);
    }
    ~MOIHGP() { if (_h) moihgp_cuda_destroy(_h); }

    // moihgp.h:148-226  step(x, y, dx, xnew, yhat, dxnew)
    void step(const State& x, const Vec& y, const DState& dx, State& xnew, Vec& yhat, DState& dxnew) {
        pack_x(x, &_xb[0]); pack_dx(dx, &_dxb[0]); pack_y(y, &_yb[0]);
        gp32_step1(_h, &_xb[0], &_yb[0], &_dxb[0], &_xb[nx()], &_yb[_num_output], &_dxb[ndx()]);
        unpack_x(&_xb[nx()], xnew); unpack_dx(&_dxb[ndx()], dxnew); unpack_y(&_yb[_num_output], yhat);
    }
    // moihgp.h:229-301  step(x, y, dx, xnew, dxnew)
    void step(const State& x, const Vec& y, const DState& dx, State& xnew, DState& dxnew) {
        pack_x(x, &_xb[0]); pack_dx(dx, &_dxb[0]); pack_y(y, &_yb[0]);
        gp32_step2(_h, &_xb[0], &_yb[0], &_dxb[0], &_xb[nx()], &_dxb[ndx()]);
        unpack_x(&_xb[nx()], xnew); unpack_dx(&_dxb[ndx()], dxnew);
    }
    // moihgp.h:304-378  step(x, y, xnew, yhat)
    void step(const State& x, const Vec& y, State& xnew, Vec& yhat) {
        pack_x(x, &_xb[0]); pack_y(y, &_yb[0]);
        gp32_step3(_h, &_xb[0], &_yb[0], &_xb[nx()], &_yb[_num_output]);
        unpack_x(&_xb[nx()], xnew); unpack_y(&_yb[_num_output], yhat);
    }
    // moihgp.h:381-428  step(x, xnew, yhat)   (prediction only)
    void step(const State& x, State& xnew, Vec& yhat) {
        pack_x(x, &_xb[0]);
        gp32_step4(_h, &_xb[0], &_xb[nx()], &_yb[_num_output]);
        unpack_x(&_xb[nx()], xnew); unpack_y(&_yb[_num_output], yhat);
    }

    // moihgp.h:431-457
    void update(const Vec& params) {
        for (size_t i = 0; i < _num_param; ++i) _pb[i] = params[i];
        detail::check(_h, moihgp_cuda_update(_h, &_pb[0]), "MOIHGP::update");
    }
    // moihgp.h:460-611
    double negLogLikelihood(const State& x, const Vec& y, const DState& dx, Vec& grad) {
        pack_x(x, &_xb[0]); pack_dx(dx, &_dxb[0]); pack_y(y, &_yb[0]);
        const double loss = gp32_lik1(_h, &_xb[0], &_yb[0], &_dxb[0], &_pb[0]);
        grad.resize(_num_param);
        for (size_t i = 0; i < _num_param; ++i) grad[i] = _pb[i];
        return loss;
    }
    // moihgp.h:614-688
    double negLogLikelihood(const State& x, const Vec& y) {
        pack_x(x, &_xb[0]); pack_y(y, &_yb[0]);
        return gp32_lik2(_h, &_xb[0], &_yb[0]);
    }

    size_t getIGPDim() { return _dim; }                  // moihgp.h:691
    size_t getNumOutput() { return _num_output; }        // moihgp.h:697
    size_t getNumLatent() { return _num_latent; }        // moihgp.h:703
    size_t getNumParam() { return _num_param; }          // moihgp.h:709
    size_t getNumIGPParam() { return _igp_num_param; }   // moihgp.h:715
    Vec getParams() {                                    // moihgp.h:721-738
        moihgp_cuda_get_params(_h, &_pb[0]);
        Vec out = detail::make_vec<Vec>(_num_param);
        for (size_t i = 0; i < _num_param; ++i) out[i] = _pb[i];
        return out;
    }

    // ---- whole-sequence entry points (the device boundary moved up to the callers' loops) ----------------------------
    // sum_t negLogLikelihood(x_t, y_t, dx_t, g) with step(x, y, dx, xnew, dxnew) advancing the state, over the T rows of
    // the packed Y[T][p]; x/dx are the carried-in state and receive the final state.
    double objective(const double* Y, size_t T, State& x, DState& dx, Vec& grad) {
        pack_x(x, &_xb[0]); pack_dx(dx, &_dxb[0]);
        double loss = 0.0;
        detail::check(_h, moihgp_cuda_objective(_h, Y, 1, T, &_xb[0], &_dxb[0], &loss, &_pb[0], &_xb[nx()], &_dxb[ndx()]), "MOIHGP::objective");
        unpack_x(&_xb[nx()], x); unpack_dx(&_dxb[ndx()], dx);
        grad.resize(_num_param);
        for (size_t i = 0; i < _num_param; ++i) grad[i] = _pb[i];
        return loss;
    }
    // loop of step(x, y, xnew, yhat) (MOIHGPRegression::predict, moihgp_regression.h:127-139) over packed Y[T][p];
    // Yhat[T][p]; optionally the filtered / smoothed states X, Xs [T][L][d] and the summed negLogLikelihood(x, y).
    void predict(const double* Y, size_t T, State& x, double* Yhat, double* X = NULL, double* Xs = NULL,
                 int smoother_mode = MOIHGP_SMOOTH_NONE, double* nll = NULL) {
        pack_x(x, &_xb[0]);
        detail::check(_h, moihgp_cuda_filter_smoother_nll(_h, Y, 1, T, &_xb[0], smoother_mode, X, Xs, Yhat, nll, &_xb[nx()]), "MOIHGP::predict");
        unpack_x(&_xb[nx()], x);
    }

    // Keep the data set resident on the device across objective evaluations (the optimiser calls the functor tens of times
    // on the same observations); the caller re-binds after changing them.  Y = NULL releases it.
    void bindData(const double* Y, size_t T) { detail::check(_h, moihgp_cuda_bind_data(_h, Y, Y ? 1 : 0, T), "MOIHGP::bindData"); }
    double objectiveBound(State& x, DState& dx, Vec& grad) {
        pack_x(x, &_xb[0]); pack_dx(dx, &_dxb[0]);
        double loss = 0.0;
        detail::check(_h, moihgp_cuda_objective_bound(_h, &_xb[0], &_dxb[0], &loss, &_pb[0], &_xb[nx()], &_dxb[ndx()]), "MOIHGP::objectiveBound");
        unpack_x(&_xb[nx()], x); unpack_dx(&_dxb[ndx()], dx);
        grad.resize(_num_param);
        for (size_t i = 0; i < _num_param; ++i) grad[i] = _pb[i];
        return loss;
    }

    // One long sequence sharded in TIME over several devices (one MOIHGP object per device / process): the
    // carry algebra a caller needs around moihgp_cuda_objective_begin_dev / _finish_dev and moihgp_cuda_fsn_block_dev.
    //   blockTransition(n, out): out[L][4][d*d] = [AKHA^n, E_0(n), E_1(n), E_2(n)], row-major
    //   smootherPower(mode, n, out): out[L][d*d] = G[mode]^n
    void blockTransition(size_t n, std::vector<double>& out) {
        out.assign(_num_latent * 4 * _dim * _dim, 0.0);
        detail::check(_h, moihgp_cuda_block_transition(_h, n, &out[0]), "MOIHGP::blockTransition");
    }
    void smootherPower(int smoother_mode, size_t n, std::vector<double>& out) {
        out.assign(_num_latent * _dim * _dim, 0.0);
        detail::check(_h, moihgp_cuda_smoother_power(_h, smoother_mode, n, &out[0]), "MOIHGP::smootherPower");
    }

    moihgp_handle* handle() { return _h; }

private:
    MOIHGP(const MOIHGP&);
    MOIHGP& operator=(const MOIHGP&);
    size_t nx() const { return _num_latent * _dim; }
    size_t ndx() const { return _num_latent * _igp_num_param * _dim; }
    void pack_x(const State& x, double* b) const { for (size_t l = 0; l < _num_latent; ++l) for (size_t i = 0; i < _dim; ++i) b[l * _dim + i] = x[l][i]; }
    void unpack_x(const double* b, State& x) const {
        x.resize(_num_latent);
        for (size_t l = 0; l < _num_latent; ++l) { x[l].resize(_dim); for (size_t i = 0; i < _dim; ++i) x[l][i] = b[l * _dim + i]; }
    }
    void pack_dx(const DState& dx, double* b) const {
        for (size_t l = 0; l < _num_latent; ++l) for (size_t k = 0; k < _igp_num_param; ++k) for (size_t i = 0; i < _dim; ++i)
            b[(l * _igp_num_param + k) * _dim + i] = dx[l][k][i];
    }
    void unpack_dx(const double* b, DState& dx) const {
        dx.resize(_num_latent);
        for (size_t l = 0; l < _num_latent; ++l) {
            dx[l].resize(_igp_num_param);
            for (size_t k = 0; k < _igp_num_param; ++k) { dx[l][k].resize(_dim); for (size_t i = 0; i < _dim; ++i) dx[l][k][i] = b[(l * _igp_num_param + k) * _dim + i]; }
        }
    }
    void pack_y(const Vec& y, double* b) const { for (size_t i = 0; i < _num_output; ++i) b[i] = y[i]; }
    void unpack_y(const double* b, Vec& y) const { y.resize(_num_output); for (size_t i = 0; i < _num_output; ++i) y[i] = b[i]; }

    moihgp_handle* _h;
    size_t _num_output, _num_latent, _dim, _num_param, _igp_num_param;
    std::vector<double> _xb, _dxb, _yb, _pb;
};

// ------------------------------------------------------------------------------------------------------------------
// RegressionObjective<StateSpace>   (moihgp_regression.h:17-70): the L-BFGS-B functor over the data set Y.
// The reference's operator() never calls _gp->update(params) (it optimises a function that is constant in
// params); `update_params = false` keeps that, `true` evaluates the objective AT params like OnlineObjective does.
template <typename StateSpace, typename Vec = std::vector<double> >
class RegressionObjective {
public:
    typedef MOIHGP<StateSpace, Vec> GP;
    RegressionObjective(const size_t& num_data, GP* gp, bool update_params = false) : _gp(gp), _update(update_params), _bound(0) { Y.reserve(num_data); }

    // Copy Y to the device once: every later operator() evaluates on that resident copy until rebind() / unbind().
    // (The optimiser calls the functor tens of times on the same data; call rebind() after changing Y.)
    void rebind() {
        const size_t p = _gp->getNumOutput(), T = Y.size();
        _buf.resize(T * p);
        for (size_t t = 0; t < T; ++t) for (size_t r = 0; r < p; ++r) _buf[t * p + r] = Y[t][r];
        _gp->bindData(T ? &_buf[0] : NULL, T);
        _bound = T;
    }
    void unbind() { _gp->bindData(NULL, 0); _bound = 0; }

    double operator()(const Vec& params, Vec& grad) {
        if (_update) _gp->update(params);
        if (_bound) {
            typename GP::State xb(_gp->getNumLatent(), detail::make_vec<Vec>(_gp->getIGPDim()));
            typename GP::DState dxb(_gp->getNumLatent(), std::vector<Vec>(_gp->getNumIGPParam(), detail::make_vec<Vec>(_gp->getIGPDim())));
            return _gp->objectiveBound(xb, dxb, grad);
        }
        const size_t p = _gp->getNumOutput(), T = Y.size();
        _buf.resize(T * p);
        for (size_t t = 0; t < T; ++t) for (size_t r = 0; r < p; ++r) _buf[t * p + r] = Y[t][r];
        typename GP::State x(_gp->getNumLatent(), detail::make_vec<Vec>(_gp->getIGPDim()));                 // zero start, :38-41
        typename GP::DState dx(_gp->getNumLatent(), std::vector<Vec>(_gp->getNumIGPParam(), detail::make_vec<Vec>(_gp->getIGPDim())));
        if (T == 0) { grad.resize(_gp->getNumParam()); for (size_t i = 0; i < _gp->getNumParam(); ++i) grad[i] = 0.0; return 0.0; }
        return _gp->objective(&_buf[0], T, x, dx, grad);                                                    // :42-50 in one pass
    }

    std::vector<Vec> Y;                                  // moihgp_regression.h:55

private:
    GP* _gp;
    bool _update;
    size_t _bound;
    std::vector<double> _buf;
};

// A stand-in for LBFGSpp::BFGSMat<double, true> with no corrections stored (get_m() == 0): the proximal term of
// OnlineObjective then uses the identity (moihgp_online.h:51-54).  Any type with get_m() and
// apply_Hv(v, a, res) -- e.g. the reference's vendored LBFGSpp::BFGSMat -- can be used instead.
template <typename Vec>
struct NoBFGSMat {
    int get_m() const { return 0; }
    void apply_Hv(const Vec&, const double&, Vec&) const {}
};

// ------------------------------------------------------------------------------------------------------------------
// OnlineObjective<StateSpace>   (moihgp_online.h:18-115): sliding-window objective with the BFGS-proximal term.
template <typename StateSpace, typename Vec = std::vector<double>, typename BFGSMat = NoBFGSMat<Vec> >
class OnlineObjective {
public:
    typedef MOIHGP<StateSpace, Vec> GP;
    OnlineObjective(GP* gp, const double& gamma, const size_t& windowsize) : _gp(gp), _gamma(gamma), _windowsize(windowsize) {
        oldparams = _gp->getParams();
        _x = typename GP::State(_gp->getNumLatent(), detail::make_vec<Vec>(_gp->getIGPDim()));
        _dx = typename GP::DState(_gp->getNumLatent(), std::vector<Vec>(_gp->getNumIGPParam(), detail::make_vec<Vec>(_gp->getIGPDim())));
        ma = detail::make_vec<Vec>(_gp->getNumOutput());
    }

    // moihgp_online.h:40-72
    double operator()(const Vec& params, Vec& grad) {
        const size_t np = _gp->getNumParam(), p = _gp->getNumOutput();
        Vec dparams = detail::make_vec<Vec>(np), Bp = detail::make_vec<Vec>(np);
        for (size_t i = 0; i < np; ++i) dparams[i] = params[i] - oldparams[i];
        _gp->update(params);                                                     // :43
        if (bfgs_mat.get_m() > 0) bfgs_mat.apply_Hv(dparams, _gamma, Bp);        // :45-48
        else for (size_t i = 0; i < np; ++i) Bp[i] = dparams[i];                 // :50-53
        double loss = 0.0;
        for (size_t i = 0; i < np; ++i) loss += dparams[i] * Bp[i];
        loss *= 0.5;                                                             // :54
        const size_t W = Y.size();
        _buf.resize(W * p);
        size_t t = 0;
        for (typename std::list<Vec>::const_iterator it = Y.begin(); it != Y.end(); ++it, ++t)
            for (size_t r = 0; r < p; ++r) _buf[t * p + r] = (*it)[r] - ma[r];   // :63
        typename GP::State x = _x;                                               // :57-58
        typename GP::DState dx = _dx;
        Vec g = detail::make_vec<Vec>(np);
        if (W > 0) loss += _gp->objective(&_buf[0], W, x, dx, g);                // :61-70 in one pass
        grad.resize(np);
        for (size_t i = 0; i < np; ++i) grad[i] = Bp[i] + g[i];
        return loss;
    }

    // moihgp_online.h:75-93 (including Q12: the carried state advances with the NEW front element)
    void push_back(const Vec& y) {
        const size_t p = _gp->getNumOutput();
        Y.push_back(y);
        for (size_t r = 0; r < p; ++r) ma[r] = 0.0;
        for (typename std::list<Vec>::const_iterator it = Y.begin(); it != Y.end(); ++it) for (size_t r = 0; r < p; ++r) ma[r] += (*it)[r];
        for (size_t r = 0; r < p; ++r) ma[r] /= double(Y.size());
        while (Y.size() > _windowsize) {
            Y.pop_front();
            Vec yc = detail::make_vec<Vec>(p);
            for (size_t r = 0; r < p; ++r) yc[r] = Y.front()[r] - ma[r];
            typename GP::State xnew;
            typename GP::DState dxnew;
            _gp->step(_x, yc, _dx, xnew, dxnew);
            _x = xnew;
            _dx = dxnew;
        }
    }

    Vec oldparams;                                       // moihgp_online.h:96-99
    BFGSMat bfgs_mat;
    std::list<Vec> Y;
    Vec ma;

private:
    GP* _gp;
    double _gamma;
    size_t _windowsize;
    typename GP::State _x;
    typename GP::DState _dx;
    std::vector<double> _buf;
};

}  // namespace moihgp_b200
#endif  // MOIHGP_B200_MOIHGP_HPP
