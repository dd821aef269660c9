// DROP-IN for moihgp/include/moihgp/matern32ss.h: the state-space arithmetic runs in the K-setup kernel; the class is the tag.
#ifndef MOIHGP_B200_DROPIN_MATERN32SS_H
#define MOIHGP_B200_DROPIN_MATERN32SS_H
#include "../../moihgp.hpp"
namespace moihgp { typedef moihgp_b200::Matern32StateSpace Matern32StateSpace; }   // matern32ss.h:13
#endif
