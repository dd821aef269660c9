// DROP-IN for moihgp/include/moihgp/moihgp.h: same include path, same names in namespace moihgp, Eigen types - backed by
// libmoihgp.so (B200).  Put include/moihgp_b200/dropin BEFORE the reference's include directory on the include path
// (the reference's include/ is still needed for its vendored LBFGSpp/), and link -lmoihgp.
#ifndef MOIHGP_B200_DROPIN_MOIHGP_H
#define MOIHGP_B200_DROPIN_MOIHGP_H
#include <Eigen/Core>
#include "../../moihgp.hpp"
#include "ihgp.h"
namespace moihgp {
template <typename StateSpace> using MOIHGP = moihgp_b200::MOIHGP<StateSpace, Eigen::VectorXd, Eigen::MatrixXd>;   // moihgp.h:76
}
#endif
