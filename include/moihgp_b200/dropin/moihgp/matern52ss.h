// DROP-IN for moihgp/include/moihgp/matern52ss.h: the state-space arithmetic runs in the K-setup kernel; the class is the tag.
#ifndef MOIHGP_B200_DROPIN_MATERN52SS_H
#define MOIHGP_B200_DROPIN_MATERN52SS_H
#include "../../moihgp.hpp"
namespace moihgp { typedef moihgp_b200::Matern52StateSpace Matern52StateSpace; }   // matern52ss.h:13
#endif
