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

// the instantiated shape that serves a model with p outputs, L latents and state dimension dim on sequences of T steps:
// (p, L) itself, or the smallest instantiated (P >= p, Lt >= L) - padded variants of the kernels: zero columns of Y / zero
// rows of U for the outputs, idle lanes / zero columns of U for the latents.  Padded latents need Lt >= 2 and 16-byte
// aligned runs of X: L * dim even, or - L * dim odd - an even T (then every sequence, round and ragged tail starts and ends
// on a 16-byte boundary all the same).  L = 1 with dim = 3 is not served (a sequence-round of X is 3 doubles).
static const Shape* find_shape(int p, int L, int dim, long long T) {
    if (L == 1 && dim == 3) return nullptr;
    const bool runs_ok = (L * dim) % 2 == 0 || (T > 0 && T % 2 == 0);
    const Shape* best = nullptr;
    for (const Shape& s : kShapes) {
        if (s.L < L || s.p < p || p < L) continue;
        if (s.L != L && (s.L < 2 || L < 2 || !runs_ok)) continue;
        if (!best || s.L < best->L || (s.L == best->L && s.p < best->p)) best = &s;
    }
    return best;
}

bool chain_supported(int p, int L, int dim, long long T) {
    if (dim != 2 && dim != 3) return false;
    return find_shape(p, L, dim, T) != nullptr;
}

// the automatic path choice: every served shape but (P = 32, L = 16) and padded latents under P = 32, where the chunked-scan
// path ties or wins (profiles/r02/chain_vs_scan_by_shape_v2_square_fix.txt, chain_vs_scan_padded_p.txt,
// chain_vs_scan_padded_L.txt: 0.69 - 1.05x)
bool chain_preferred(int p, int L, int dim, long long T) {
    if (dim != 2 && dim != 3) return false;
    const Shape* s = find_shape(p, L, dim, T);
    return s != nullptr && !(s->p == 32 && (s->L == 16 || s->L != L));
}

cudaError_t launch_chain(int p, int L, int dim, const ChainArgs& a, cudaStream_t st) {
    const Shape* s = find_shape(p, L, dim, a.T);
    if (!s) return cudaErrorInvalidValue;
    ChainArgs b = a;
    b.p = p;
    b.L = L;
    return s->fn(dim, b, st);
}

}  // namespace moihgp
