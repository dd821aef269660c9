// Compile-only check: the host mirror instantiates with an Eigen::VectorXd-shaped vector type (here the Eigen-API shim of
// oracle/eigen_shim, test infrastructure), i.e. with the reference's exact functor signature
//   double operator()(const Eigen::VectorXd& params, Eigen::VectorXd& grad)      (moihgp_regression.h:34, moihgp_online.h:40)
#include <Eigen/Core>
#include <moihgp_b200/moihgp.hpp>

namespace moihgp = moihgp_b200;   // the one-line switch INTEGRATION.md describes

double use(moihgp::MOIHGP<moihgp::Matern32StateSpace, Eigen::VectorXd>* gp, const Eigen::VectorXd& params, Eigen::VectorXd& grad) {
    moihgp::RegressionObjective<moihgp::Matern32StateSpace, Eigen::VectorXd> f(63, gp, true);
    moihgp::OnlineObjective<moihgp::Matern52StateSpace, Eigen::VectorXd>* fo = 0;
    (void)fo;
    return f(params, grad);
}
