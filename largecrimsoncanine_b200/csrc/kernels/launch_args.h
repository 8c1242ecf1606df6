// launch_args.h -- plain argument blocks handed from the C-ABI dispatcher (lcc_abi.cu) to the
// per-signature kernel translation units.  Internal; nothing here crosses the C ABI.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lcc {

enum : int { kSoA = 0, kAoS = 1 };
enum : int { kF32 = 0, kF64 = 1 };
// product selector inside kernels: the three reference products + the fused dual-output form
// (values 0-2 and 4-7 are the public lcc_product_op codes; 3 is internal)
enum : int { kOpGP = 0, kOpOuter = 1, kOpLC = 2, kOpGPOuter = 3, kOpRC = 4, kOpScalar = 5, kOpCommutator = 6, kOpAnticommutator = 7 };

struct ProductArgs {
    int op = kOpGP;
    int dtype = kF64;
    int layout = kSoA;
    bool strict = false;
    int64_t n = 0;  // rows of the result
    const void *a = nullptr; int64_t lda = 0; bool a_bcast = false;
    const void *b = nullptr; int64_t ldb = 0; bool b_bcast = false;
    void *out = nullptr; int64_t ldo = 0;
    void *out2 = nullptr; int64_t ldo2 = 0;  // outer-product output of the fused form
    const double *mu = nullptr;              // run-time metric factors mu[i & j] (generic kernels only)
    // fused reduction epilogue (K5 as the epilogue of K1): when sum_partial is set nothing is stored; each block
    // writes its column sums to sum_partial[blade * sum_blocks + block] (f64).  a_grade >= 0 additionally applies
    // the grade projection <a>_k to the left operand on load (masked blade rows are never read).
    double *sum_partial = nullptr;
    int sum_blocks = 0;
    int a_grade = -1;
    int list = 0;  // term list: 0 forward product; 1 / 2 cotangent of the left / right operand (a, b = the list's operands)
    cudaStream_t stream = nullptr;
};

// Fixed grid of the fused reduction: block b always owns the same rows and writes partial[b], so results do not
// depend on n, on timing or on which SM runs a block.  Many more blocks than resident slots (148 SMs x 6 at 138
// registers and 64 threads) on purpose: (1) the tail -- the last, partly filled wave cannot saturate HBM -- shrinks
// with the block count: 4736 blocks 6775 GB/s, 8288 6989, 16576 7129, 33152 7143 (STA f64, 2^27 rows,
// profiles/r01_sta_fused_grid_sweep.txt); (2) when a concurrent kernel (the NCCL all-reduce of the previous step)
// holds a slot for a while, the hardware scheduler hands its share to the other slots at block granularity instead
// of one late block finishing 0.2 ms after everybody else (2 GPUs: 3.79 ms per step with one wave of 592 blocks).
// LCC_FUSED_SUM_BLOCKS overrides it for sweeps (lcc_abi.cu).
constexpr int kFusedSumBlocks = 148 * 112;

struct SandwichArgs {
    int dtype = kF64;
    int layout = kSoA;
    bool strict = false;
    bool even = false;  // rotor has no odd-grade part: odd blades are pruned at compile time
    int64_t n = 0;
    const double *rotor = nullptr;  // host, num_blades (one rotor for the whole batch, kernel parameters)
    const void *rotors = nullptr; int64_t ldr = 0; bool r_bcast = false;  // or: device rotors, one per row (r_bcast: one row for all)
    const void *x = nullptr; int64_t ldx = 0;
    void *out = nullptr; int64_t ldo = 0;
    // fused cotangent pass of the per-row sandwich (rotors set): grad = d/d out -> gx = d/dx, gr = d/d rotor rows
    const void *grad = nullptr; int64_t ldg = 0;
    void *gx = nullptr; int64_t ldgx = 0;
    void *gr = nullptr; int64_t ldgr = 0;
    const double *mu = nullptr;
    cudaStream_t stream = nullptr;
};

// Each generated translation unit exports one pair of launchers with this shape; they return
// cudaSuccess, a launch error, or cudaErrorNotSupported when the (dtype, layout) is not built.
using ProductLauncher = cudaError_t (*)(const ProductArgs &);
using SandwichLauncher = cudaError_t (*)(const SandwichArgs &);

// tuning options (lcc_set_option): thresholds of the TMA-staged kernels
int64_t option_aos_tma_min_rows();
int64_t option_soa_tma_min_bytes();

// counted by every launcher (bench.py reports it as gpu_launches)
void note_kernel_launch();

}  // namespace lcc
