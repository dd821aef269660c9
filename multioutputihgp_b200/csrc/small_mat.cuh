// d x d (d = 2, 3) matrix-vector helpers shared by the scan and objective kernels.  Matrices in LatentConsts are
// row-major with a fixed row stride of 3 (moihgp_device.cuh); in registers they are packed d x d.
#pragma once

namespace moihgp {

template <int D> __device__ __forceinline__ void load_mat(const double* src9, double* dst) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) dst[i * D + j] = __ldg(src9 + i * 3 + j);
}
template <int D> __device__ __forceinline__ void load_vec(const double* src3, double* dst) {
#pragma unroll
    for (int i = 0; i < D; ++i) dst[i] = __ldg(src3 + i);
}
template <int D> __device__ __forceinline__ void mv(const double* M, const double* x, double* y) {  // y = M x
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = M[i * D] * x[0];
#pragma unroll
        for (int j = 1; j < D; ++j) s = fma(M[i * D + j], x[j], s);
        y[i] = s;
    }
}
template <int D> __device__ __forceinline__ void mv_acc(const double* M, const double* x, double* y) {  // y += M x
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = y[i];
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(M[i * D + j], x[j], s);
        y[i] = s;
    }
}

}  // namespace moihgp
