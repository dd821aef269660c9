// many-chains kernels for p = 4 outputs, L = 1 latent (Matern-3/2 only: see chain.cu) (see chain_kernels.cuh)
#include "chain_kernels.cuh"
MOIHGP_CHAIN_INSTANCE(4, 1, false)
