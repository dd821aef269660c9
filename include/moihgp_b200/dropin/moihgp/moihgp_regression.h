// DROP-IN for moihgp/include/moihgp/moihgp_regression.h: RegressionObjective and MOIHGPRegression in namespace moihgp, driven
// by the reference's own vendored LBFGS++ (<LBFGSpp/LBFGSB.h>, found on the reference's include path).
#ifndef MOIHGP_B200_DROPIN_MOIHGP_REGRESSION_H
#define MOIHGP_B200_DROPIN_MOIHGP_REGRESSION_H
#include <Eigen/Core>
#include <LBFGSpp/LBFGSB.h>
#include "moihgp.h"
#include "../../learners.hpp"
namespace moihgp {
template <typename StateSpace> using RegressionObjective = moihgp_b200::RegressionObjective<StateSpace, Eigen::VectorXd, Eigen::MatrixXd>;   // :17
template <typename StateSpace> using MOIHGPRegression =                                                                                     // :73
    moihgp_b200::MOIHGPRegression<StateSpace, Eigen::VectorXd, Eigen::MatrixXd, LBFGSpp::LBFGSBSolver<double>, LBFGSpp::LBFGSBParam<double> >;
}
#endif
