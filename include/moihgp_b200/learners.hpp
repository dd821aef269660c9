// moihgp_b200/learners.hpp - the reference's two learner shells around LBFGS++, over the B200 objective functors:
//   moihgp::MOIHGPRegression<StateSpace>      moihgp/include/moihgp/moihgp_regression.h:73-202
//   moihgp::MOIHGPOnlineLearning<StateSpace>  moihgp/include/moihgp/moihgp_online.h:118-255
// Same constructor arguments, box bounds, L-BFGS-B settings, method names and public members as the reference.  The
// optimiser itself is NOT part of this library: the reference vendors (and patches) LBFGS++ and keeps it on the host
// (SURVEY 2 row 10), so the solver and its parameter struct are template parameters -
//   MOIHGPRegression<SS, Eigen::VectorXd, Eigen::MatrixXd, LBFGSpp::LBFGSBSolver<double>, LBFGSpp::LBFGSBParam<double> >
// - and `_solver->minimize(*_obj, _params, fx, _lb, _ub)` (moihgp_regression.h:121, moihgp_online.h:185; LBFGSB.h:116-241)
// drives the device objective exactly as it drives the reference's CPU functor.  dropin/moihgp/moihgp_regression.h and
// dropin/moihgp/moihgp_online.h make these the `moihgp::` classes under the reference's own include paths.
#ifndef MOIHGP_B200_LEARNERS_HPP
#define MOIHGP_B200_LEARNERS_HPP

#include "moihgp.hpp"

namespace moihgp_b200 {

namespace detail {
// box bounds of both learners (moihgp_regression.h:91-98, moihgp_online.h:133-140)
template <typename Vec>
inline void learner_bounds(size_t p, size_t L, size_t igp_num_param, size_t num_param, Vec& lb, Vec& ub) {
    lb = make_vec<Vec>(num_param);
    ub = make_vec<Vec>(num_param);
    for (size_t i = 0; i < p * L; ++i) { lb[i] = -1e+4; ub[i] = 1e+4; }                       // U block
    for (size_t i = p * L; i < p * L + L; ++i) { lb[i] = 1e-4; ub[i] = 1e+4; }               // S
    for (size_t i = num_param - (igp_num_param * L + 1); i < num_param; ++i) { lb[i] = 1e-4; ub[i] = 1e+2; }   // sigma, IGP parameters
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------------------------
template <typename StateSpace, typename Vec, typename Mat, typename Solver, typename SolverParam>
class MOIHGPRegression {
public:
    typedef MOIHGP<StateSpace, Vec, Mat> GP;
    typedef RegressionObjective<StateSpace, Vec, Mat> Objective;

    // moihgp_regression.h:80-108
    MOIHGPRegression(const double& dt, const size_t& num_output, const size_t& num_latent, const size_t& num_data, const bool& threading)
        : _threading(threading), _dt(dt), _num_output(num_output), _num_latent(num_latent), _num_data(num_data) {
        _moihgp = new GP(dt, num_output, num_latent, threading);
        _dim = _moihgp->getIGPDim();
        _num_param = _moihgp->getNumParam();
        _igp_num_param = _moihgp->getNumIGPParam();
        detail::learner_bounds(_num_output, _num_latent, _igp_num_param, _num_param, _lb, _ub);
        _params = _moihgp->getParams();
        _LBFGSB_param.max_iterations = 1000;
        _LBFGSB_param.m = 10;
        _LBFGSB_param.max_linesearch = 20;
        _LBFGSB_param.ftol = 1e-8;
        _LBFGSB_param.epsilon = 1e-8;
        _LBFGSB_param.epsilon_rel = 1e-8;
        _solver = new Solver(_LBFGSB_param);
        _obj = new Objective(_num_data, _moihgp);        // as the reference: the functor does NOT call update(params)
    }
    ~MOIHGPRegression() {
        delete _obj;
        delete _solver;
        delete _moihgp;
    }

    // moihgp_regression.h:118-124.  The data set goes to the device once; every evaluation of the line search is then one
    // fused device pass on the resident copy.
    int fit(const std::vector<Vec>& Y) {
        _obj->Y = Y;
        _obj->rebind();
        double fx;
        int num_iter = _solver->minimize(*_obj, _params, fx, _lb, _ub);
        _obj->unbind();
        _params = _moihgp->getParams();
        return num_iter;
    }

    // moihgp_regression.h:127-139: the loop of step(x, y, xnew, yhat) from a zero state, as one device pass
    std::vector<Vec> predict(const std::vector<Vec>& Y) {
        const size_t T = Y.size(), p = _num_output;
        std::vector<Vec> Yhat;
        Yhat.reserve(T);
        if (T == 0) return Yhat;
        std::vector<double> yin(T * p), yout(T * p);
        for (size_t t = 0; t < T; ++t) for (size_t r = 0; r < p; ++r) yin[t * p + r] = Y[t][r];
        typename GP::State x(_num_latent, detail::make_vec<Vec>(_dim));
        _moihgp->predict(&yin[0], T, x, &yout[0]);
        for (size_t t = 0; t < T; ++t) {
            Vec yh = detail::make_vec<Vec>(p);
            for (size_t r = 0; r < p; ++r) yh[r] = yout[t * p + r];
            Yhat.push_back(yh);
        }
        return Yhat;
    }

    Vec getParams() { return _moihgp->getParams(); }     // moihgp_regression.h:142-146
    size_t getNumParam() { return _num_param; }
    size_t getNumOutput() { return _num_output; }
    size_t getNumLatent() { return _num_latent; }
    size_t getNumIGPParam() { return _igp_num_param; }
    size_t getIGPDim() { return _dim; }
    size_t getNumData() { return _num_data; }

private:
    MOIHGPRegression(const MOIHGPRegression&);
    MOIHGPRegression& operator=(const MOIHGPRegression&);
    GP* _moihgp;
    double _threading;
    double _dt;
    size_t _num_output, _num_latent, _num_data, _num_param, _igp_num_param, _dim;
    Vec _params;
    SolverParam _LBFGSB_param;
    Solver* _solver;
    Vec _lb, _ub;
    Objective* _obj;
};

// ------------------------------------------------------------------------------------------------------------------
template <typename StateSpace, typename Vec, typename Mat, typename Solver, typename SolverParam, typename BFGSMat>
class MOIHGPOnlineLearning {
public:
    typedef MOIHGP<StateSpace, Vec, Mat> GP;
    typedef OnlineObjective<StateSpace, Vec, BFGSMat, Mat> Objective;

    // moihgp_online.h:123-162
    MOIHGPOnlineLearning(const double& dt, const size_t& num_output, const size_t& num_latent, const double& gamma, const size_t& windowsize,
                         const bool& threading)
        : _threading(threading), _dt(dt), _num_output(num_output), _num_latent(num_latent), _gamma(gamma) {
        _moihgp = new GP(dt, num_output, num_latent, threading);
        _dim = _moihgp->getIGPDim();
        _igp_num_param = _moihgp->getNumIGPParam();
        _num_param = _moihgp->getNumParam();
        detail::learner_bounds(_num_output, _num_latent, _igp_num_param, _num_param, _lb, _ub);
        x.assign(_num_latent, detail::make_vec<Vec>(_dim));
        dx.assign(_num_latent, std::vector<Vec>(_igp_num_param, detail::make_vec<Vec>(_dim)));
        _windowsize = windowsize < 1 ? 1 : windowsize;
        _params = _moihgp->getParams();
        _LBFGSB_param.m = 10;
        _LBFGSB_param.max_iterations = 5;
        _LBFGSB_param.max_linesearch = 20;
        _LBFGSB_param.max_step = 1e-1;
        _LBFGSB_param.ftol = 1e-8;
        _LBFGSB_param.epsilon = 1e-8;
        _LBFGSB_param.epsilon_rel = 1e-8;
        _solver = new Solver(_LBFGSB_param);
        _obj = new Objective(_moihgp, _gamma, _windowsize);
    }
    ~MOIHGPOnlineLearning() {
        delete _obj;
        delete _solver;
        delete _moihgp;
    }

    // moihgp_online.h:173-187
    Vec step(const Vec& y) {
        Vec yhat, yc = detail::make_vec<Vec>(_num_output);
        typename GP::State xnew;
        typename GP::DState dxnew(_num_latent, std::vector<Vec>(_igp_num_param, detail::make_vec<Vec>(_dim)));
        _obj->push_back(y);
        for (size_t r = 0; r < _num_output; ++r) yc[r] = y[r] - _obj->ma[r];
        _moihgp->step(x, yc, xnew, yhat);
        for (size_t r = 0; r < _num_output; ++r) yhat[r] += _obj->ma[r];
        x = xnew;
        dx = dxnew;                                      // zeros, as the reference (:181)
        _obj->bfgs_mat = _solver->getBFGSMat();
        _obj->oldparams = _params;
        _obj->prepare();                                 // the proximal term goes to the device once per sample
        double fx;
        _solver->minimize(*_obj, _params, fx, _lb, _ub);
        return yhat;
    }

    Vec getParams() { return _moihgp->getParams(); }     // moihgp_online.h:190-194
    size_t getNumParam() { return _num_param; }
    size_t getNumOutput() { return _num_output; }
    size_t getNumLatent() { return _num_latent; }
    size_t getNumIGPParam() { return _igp_num_param; }
    size_t getIGPDim() { return _dim; }
    size_t getWindowsize() { return _windowsize; }

    typename GP::State x;                                // moihgp_online.h:232-233
    typename GP::DState dx;

private:
    MOIHGPOnlineLearning(const MOIHGPOnlineLearning&);
    MOIHGPOnlineLearning& operator=(const MOIHGPOnlineLearning&);
    GP* _moihgp;
    bool _threading;
    double _dt;
    size_t _dim, _num_output, _num_latent, _num_param, _igp_num_param, _windowsize;
    double _gamma;
    Vec _params;
    SolverParam _LBFGSB_param;
    Solver* _solver;
    Vec _lb, _ub;
    Objective* _obj;
};

}  // namespace moihgp_b200
#endif  // MOIHGP_B200_LEARNERS_HPP
