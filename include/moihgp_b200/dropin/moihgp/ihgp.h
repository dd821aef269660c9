// DROP-IN for moihgp/include/moihgp/ihgp.h (see moihgp.h in this directory).
#ifndef MOIHGP_B200_DROPIN_IHGP_H
#define MOIHGP_B200_DROPIN_IHGP_H
#include <Eigen/Core>
#include "../../moihgp.hpp"
namespace moihgp {
template <typename StateSpace> using IHGP = moihgp_b200::IHGP<StateSpace, Eigen::VectorXd, Eigen::MatrixXd>;       // ihgp.h:17
}
#endif
