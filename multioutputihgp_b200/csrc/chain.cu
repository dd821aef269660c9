// Many-chains path: dispatch over the instantiated (outputs, latents) shapes.  The kernels live in chain_kernels.cuh and
// are instantiated one shape per translation unit (chain_inst_PxL.cu).
#include <cuda_runtime.h>
#include "launch.h"

namespace moihgp {

cudaError_t launch_chain_16x8(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_8x4(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_8x2(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_8x8(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_16x2(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_16x4(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_16x16(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_32x4(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_32x8(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_32x16(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_4x2(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_4x4(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_32x2(int dim, const ChainArgs& a, cudaStream_t st);

namespace {
struct Shape { int p, L; cudaError_t (*fn)(int, const ChainArgs&, cudaStream_t); };
const Shape kShapes[] = {
    {16, 8, launch_chain_16x8},
    {8, 4, launch_chain_8x4},
    {8, 2, launch_chain_8x2},
    {8, 8, launch_chain_8x8},
    {16, 2, launch_chain_16x2},
    {16, 4, launch_chain_16x4},
    {16, 16, launch_chain_16x16},
    {32, 4, launch_chain_32x4},
    {32, 8, launch_chain_32x8},
    {32, 16, launch_chain_32x16},
    {4, 2, launch_chain_4x2},
    {4, 4, launch_chain_4x4},
    {32, 2, launch_chain_32x2},
};
}  // namespace

bool chain_supported(int p, int L, int dim) {
    if (dim != 2 && dim != 3) return false;
    for (const Shape& s : kShapes) if (s.p == p && s.L == L) return true;
    return false;
}

cudaError_t launch_chain(int p, int L, int dim, const ChainArgs& a, cudaStream_t st) {
    for (const Shape& s : kShapes) if (s.p == p && s.L == L) return s.fn(dim, a, st);
    return cudaErrorInvalidValue;
}

}  // namespace moihgp
