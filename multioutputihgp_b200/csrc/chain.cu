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
cudaError_t launch_chain_4x1(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_8x1(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_16x1(int dim, const ChainArgs& a, cudaStream_t st);
cudaError_t launch_chain_32x1(int dim, const ChainArgs& a, cudaStream_t st);

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
    {4, 1, launch_chain_4x1},
    {8, 1, launch_chain_8x1},
    {16, 1, launch_chain_16x1},
    {32, 1, launch_chain_32x1},
};
}  // namespace

// the instantiated shape that serves a model with p outputs and L latents: (p, L) itself, or the next larger
// instantiated P with the same L (padded variant of k_filter_chain: zero columns, zero rows of U), else null
static const Shape* find_shape(int p, int L) {
    const Shape* best = nullptr;
    for (const Shape& s : kShapes) {
        if (s.L != L || s.p < p) continue;
        if (s.p != p && p < L) continue;
        if (!best || s.p < best->p) best = &s;
    }
    return best;
}

// L = 1 with state dimension 3: a sequence-round of X is 3 doubles, which the 16-byte copy-out of the staging tile cannot
// tile - Matern-5/2 models with a single latent take the chunked-scan path
bool chain_supported(int p, int L, int dim) {
    if (dim != 2 && dim != 3) return false;
    if (L == 1 && dim == 3) return false;
    return find_shape(p, L) != nullptr;
}

// the automatic path choice: every served shape but (P = 32, L = 16), where the chunked-scan path ties or wins
// (profiles/r02/chain_vs_scan_by_shape_v2_square_fix.txt, chain_vs_scan_padded_p.txt: 0.74 - 1.05x)
bool chain_preferred(int p, int L, int dim) {
    if (!chain_supported(p, L, dim)) return false;
    const Shape* s = find_shape(p, L);
    return s != nullptr && !(s->p == 32 && s->L == 16);
}

cudaError_t launch_chain(int p, int L, int dim, const ChainArgs& a, cudaStream_t st) {
    const Shape* s = find_shape(p, L);
    if (!s) return cudaErrorInvalidValue;
    ChainArgs b = a;
    b.p = p;
    return s->fn(dim, b, st);
}

}  // namespace moihgp
