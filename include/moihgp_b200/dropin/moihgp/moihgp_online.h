// DROP-IN for moihgp/include/moihgp/moihgp_online.h: OnlineObjective and MOIHGPOnlineLearning in namespace moihgp, driven by
// the reference's own vendored LBFGS++ (its BFGS matrix feeds the proximal term, moihgp_online.h:45-48, :182).
#ifndef MOIHGP_B200_DROPIN_MOIHGP_ONLINE_H
#define MOIHGP_B200_DROPIN_MOIHGP_ONLINE_H
#include <Eigen/Core>
#include <LBFGSpp/LBFGSB.h>
#include "moihgp.h"
#include "../../learners.hpp"
namespace moihgp {
template <typename StateSpace> using OnlineObjective =                                                                                      // :18
    moihgp_b200::OnlineObjective<StateSpace, Eigen::VectorXd, LBFGSpp::BFGSMat<double, true>, Eigen::MatrixXd>;
template <typename StateSpace> using MOIHGPOnlineLearning =                                                                                 // :118
    moihgp_b200::MOIHGPOnlineLearning<StateSpace, Eigen::VectorXd, Eigen::MatrixXd, LBFGSpp::LBFGSBSolver<double>, LBFGSpp::LBFGSBParam<double>,
                                      LBFGSpp::BFGSMat<double, true> >;
}
#endif
