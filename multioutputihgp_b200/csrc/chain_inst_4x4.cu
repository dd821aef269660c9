// many-chains kernels for p = 4 outputs, L = 4 latents (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(4, 4, false)
