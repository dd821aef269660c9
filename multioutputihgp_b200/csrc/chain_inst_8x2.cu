// many-chains kernels for p = 8 outputs, L = 2 latents (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(8, 2, false)
