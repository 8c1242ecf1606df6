// fixed_operand_matrix.cu -- the nb x nb matrix of a FIXED-operand linear map in a 64 / 128 / 256-blade algebra, built
// on the device in f64.  K6 (tcgen05, 3xTF32) and K6d (FP64 tensor pipe) then run  Out = X * Wt^T  against it.
//
//   product   x -> fixed (op) x  (side 0)  or  x -> x (op) fixed  (side 1), op in {gp, outer, left / right contraction,
//             scalar product}: the broadcast branch of /root/reference/src/batch.rs:365-404 (gp), :514-574 (outer),
//             :585-655 (lc) with the op's term mask folded into the matrix
//               side 0: Wt[n][k] = [mask(n^k, k)] sign[n^k][k] * M[n^k]        (left blade i = n^k, right blade j = k)
//               side 1: Wt[n][k] = [mask(k, n^k)] sign[k][n^k] * M[n^k]        (left blade i = k,   right blade j = n^k)
//   sandwich  x -> R x ~R with ONE rotor for the whole batch (/root/reference/src/batch.rs:452-497): the composition of a
//             left and a right multiplication is again one nb x nb matrix,
//               Wt[n][k] = sum_t  sign[n^t][t] R[n^t]  *  sign[k][k^t] R[k^t] rev(k^t)
//             so the whole sandwich is ONE GEMM instead of two products and a full-size intermediate.
// The rotor travels as a kernel parameter (2 KB), so no host staging buffer has to outlive the call: no synchronisation.
#include <cuda_runtime.h>

#include <cstdint>

#include "launch_args.h"
#include "misc_kernels.cuh"

namespace lcc {
namespace {

struct RotorParam { double v[256]; };

__device__ __forceinline__ bool in_op(int op, int i, int j) {
    return op == kOpGP || (op == kOpOuter && (i & j) == 0) || (op == kOpLC && (i & j) == i) || (op == kOpRC && (i & j) == j) ||
           (op == kOpScalar && i == j);
}

template <typename T>
__global__ void build_product_matrix_kernel(const int8_t *__restrict__ signs, const T *__restrict__ m, int64_t m_stride, int nb, int op, int side,
                                            double *__restrict__ wt) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * nb) return;
    const int n = idx / nb, k = idx % nb, other = n ^ k;
    const int i = side == 0 ? other : k, j = side == 0 ? k : other;
    const int s = in_op(op, i, j) ? signs[i * nb + j] : 0;
    const double mv = (double)m[(int64_t)other * m_stride];
    wt[idx] = s == 0 ? 0.0 : (s > 0 ? mv : -mv);
}

__global__ void build_sandwich_matrix_kernel(const int8_t *__restrict__ signs, const __grid_constant__ RotorParam R, int nb, double *__restrict__ wt) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * nb) return;
    const int n = idx / nb, k = idx % nb;
    double acc = 0.0;
    for (int t = 0; t < nb; ++t) {
        const int li = n ^ t, rj = k ^ t;  // R[li] * e_t -> e_n ;  e_k * ~R[rj] -> e_t
        const int s = (int)signs[li * nb + t] * (int)signs[k * nb + rj];
        if (s == 0) continue;
        const int g = __popc(rj);
        const double rev = (g % 4 < 2) ? 1.0 : -1.0;  // blades.rs:128-135
        acc += (double)s * rev * R.v[li] * R.v[rj];
    }
    wt[idx] = acc;
}

// rotor and reversed rotor as two one-row operands of the batch dtype (strict mode: two exact-order table products)
template <typename T>
__global__ void write_rotor_rows_kernel(const __grid_constant__ RotorParam R, int nb, T *__restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int g = __popc(i);
    const T r = (T)R.v[i];
    rows[i] = r;
    rows[nb + i] = r * ((g % 4 < 2) ? T(1) : T(-1));  // batch.rs:452-457
}

}  // namespace

cudaError_t launch_build_product_matrix(int nb, int op, int side, const int8_t *signs_dev, const void *fixed_dev, int64_t fixed_stride,
                                        int fixed_dtype, double *wt, cudaStream_t s) {
    const int blocks = (nb * nb + 255) / 256;
    if (fixed_dtype == kF32) build_product_matrix_kernel<float><<<blocks, 256, 0, s>>>(signs_dev, (const float *)fixed_dev, fixed_stride, nb, op, side, wt);
    else build_product_matrix_kernel<double><<<blocks, 256, 0, s>>>(signs_dev, (const double *)fixed_dev, fixed_stride, nb, op, side, wt);
    note_kernel_launch();
    return cudaGetLastError();
}

cudaError_t launch_build_sandwich_matrix(int nb, const int8_t *signs_dev, const double *rotor_host, double *wt, cudaStream_t s) {
    RotorParam R;
    for (int i = 0; i < 256; ++i) R.v[i] = i < nb ? rotor_host[i] : 0.0;
    build_sandwich_matrix_kernel<<<(nb * nb + 255) / 256, 256, 0, s>>>(signs_dev, R, nb, wt);
    note_kernel_launch();
    return cudaGetLastError();
}

cudaError_t launch_write_rotor_rows(int nb, int dtype, const double *rotor_host, void *rows_dev, cudaStream_t s) {
    RotorParam R;
    for (int i = 0; i < 256; ++i) R.v[i] = i < nb ? rotor_host[i] : 0.0;
    if (dtype == kF32) write_rotor_rows_kernel<float><<<1, 256, 0, s>>>(R, nb, (float *)rows_dev);
    else write_rotor_rows_kernel<double><<<1, 256, 0, s>>>(R, nb, (double *)rows_dev);
    note_kernel_launch();
    return cudaGetLastError();
}

}  // namespace lcc
