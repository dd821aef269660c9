// many-chains kernels for p = 16 outputs, L = 16 latents (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(16, 16, false)
