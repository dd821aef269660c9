// many-chains kernels for p = 8 outputs, L = 1 latent (Matern-3/2 only: see chain.cu) (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(8, 1, false)
