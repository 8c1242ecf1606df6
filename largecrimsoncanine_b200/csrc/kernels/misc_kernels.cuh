// misc_kernels.cuh -- declarations of the run-time-nb kernels' host launchers (misc_kernels.cu):
// K3 signed permutations / masks, K4 norms, K5 batch sums, K7 layout conversion, the row/column
// movers behind the container API, and the looped product used for algebras above 32 blades.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lcc {

struct ViewArg {  // one device batch; mirrors lcc_view with typed fields
    void *data; int64_t n; int64_t ld; int dtype; int layout;
};

struct BladeMap {  // out[k] = factor[k] * in[source[k]], factor in {-1, 0, +1}
    int8_t factor[256];
    uint16_t source[256];
};

cudaError_t launch_blade_map(int nb, const ViewArg &x, const ViewArg &out, const BladeMap &map, double scale,
                             cudaStream_t s);
cudaError_t launch_addsub(int nb, int op, const ViewArg &a, bool a_bcast, const ViewArg &b, bool b_bcast,
                          const ViewArg &out, cudaStream_t s);
// weights[k] = sign[k][k] * reverse_sign(k) in {-1,0,+1}; mode 0 norm^2, 1 norm, 2 normalized
cudaError_t launch_norm(int nb, int mode, bool strict, const ViewArg &x, const int8_t *weights, void *norms,
                        const ViewArg &out, cudaStream_t s);
cudaError_t launch_sum(int nb, const ViewArg &x, void *sums_dev, void *scratch_dev, int64_t scratch_elems,
                       cudaStream_t s);
int64_t sum_scratch_elems(int nb, int64_t n);
// second pass of every batch sum: sums[k] = sum over blocks of partial[k*nblocks + block] (fixed order)
cudaError_t launch_sum_final(int nb, int nblocks, const double *partial, void *sums_dev, int dtype, cudaStream_t s);
cudaError_t launch_convert(int nb, const ViewArg &src, const ViewArg &dst, cudaStream_t s);
cudaError_t launch_scatter_columns(int nb, const void *cols, int ncols, const int32_t *blade_of_col,
                                   const ViewArg &dst, cudaStream_t s);
cudaError_t launch_gather_columns(int nb, const ViewArg &src, int ncols, const int32_t *blade_of_col, void *cols,
                                  cudaStream_t s);
cudaError_t launch_copy_rows(int nb, const ViewArg &src, int64_t start, int64_t end, const ViewArg &dst,
                             int64_t dst_start, cudaStream_t s);
// Table-driven product out[k] = sum_i G[k][i]*x[i]*y[i^k] (any layout, up to 256 blades): products of
// algebras above 32 blades and the vector-Jacobian products of lcc.torch.  gsign is on the device.
cudaError_t launch_table_product(int nb, bool strict, const ViewArg &x, bool x_bcast, const ViewArg &y, bool y_bcast,
                                 const ViewArg &out, const int8_t *gsign_dev, cudaStream_t s);
// Fixed-operand linear maps in 64 / 128 / 256-blade algebras (fixed_operand_matrix.cu): the nb x nb matrix Wt (f64, device)
// of  x -> fixed (op) x (side 0) / x (op) fixed (side 1)  with the op's term mask folded in, or of  x -> R x ~R.
cudaError_t launch_build_product_matrix(int nb, int op, int side, const int8_t *signs_dev, const void *fixed_dev, int64_t fixed_stride,
                                        int fixed_dtype, double *wt, cudaStream_t s);
cudaError_t launch_build_sandwich_matrix(int nb, const int8_t *signs_dev, const double *rotor_host, double *wt, cudaStream_t s);
// rotor and ~rotor as two one-row operands (2*nb elements of `dtype`) written from kernel parameters: no host staging
cudaError_t launch_write_rotor_rows(int nb, int dtype, const double *rotor_host, void *rows_dev, cudaStream_t s);
// K6 (fixed_operand_gemm.cu): Out = X * Wt^T, f32 SoA, tcgen05 tensor cores with the 3xTF32 split; scratch_w = 2*nb*nb floats.
// K6d (fixed_operand_dmma.cu): the f64 form on the FP64 tensor pipe (mma.sync m8n8k4).
// Both return cudaErrorNotSupported for views they cannot take (pitch / alignment): the caller falls back to the table kernel.
cudaError_t launch_matrix_product_tf32x3(int nb, const double *wt64, const ViewArg &x, const ViewArg &out, float *scratch_w, cudaStream_t s);
cudaError_t launch_matrix_product_f64_dmma(int nb, const double *wt64, const ViewArg &x, const ViewArg &out, cudaStream_t s);
// K8 (matrix_rep_kernel.cu): fixed-operand geometric product (mode 0 left, 1 right) and fixed-rotor sandwich (mode 2) of
// 128 / 256-blade algebras in their 16 x 16 matrix representation: two Walsh-Hadamard passes + 16^3 FMAs per row instead
// of the nb^2 GEMM, HBM-bound.  p1 / p2: 16 x 16 f64 factors as the kernel indexes them; SoA views, n * sizeof(T) % 16 == 0.
cudaError_t launch_matrix_rep_product(int nb, int mode, const double *p1, const double *p2, const uint8_t *blade_at, const int8_t *sign_at,
                                      const ViewArg &x, const ViewArg &out, cudaStream_t s, const void *m_dev = nullptr, int64_t m_stride = 1);
// PGA3D / CGA3D point embeddings and extractors (src/pga.rs:131-134,579-599; src/cga.rs:189-206,492-508);
// which 2 / 3: STA bivectors from (E,B) and boost rotors from velocities (src/sta.rs:202-212,253-296)
cudaError_t launch_points_embed(int which, const void *xyz, const ViewArg &dst, cudaStream_t s);
cudaError_t launch_points_extract(int which, const ViewArg &src, void *xyz, cudaStream_t s);

}  // namespace lcc
