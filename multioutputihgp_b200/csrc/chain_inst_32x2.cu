// many-chains kernels for p = 32 outputs, L = 2 latents (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(32, 2, false)
