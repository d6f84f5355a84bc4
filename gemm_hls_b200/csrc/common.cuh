// Shared host-side helpers for the CUDA translation units of libmm_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/mm_b200.h"

namespace mm {

// Thread-local error message behind mm_last_error().
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define MM_CUDA_TRY(expr)                                                                  \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      return ::mm::fail(MM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    }                                                                                      \
  } while (0)

inline unsigned ceil_div(unsigned a, unsigned b) { return (a + b - 1) / b; }

// Scratch the tensor-core path needs (rounded / transposed operand copies); owned by the
// context, grown on demand, reused across calls.
struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct GemmArgs {
  const void *a;
  const void *b;
  void *c;
  unsigned n, k, m;
  int flags;
  cudaStream_t stream;
  // optional profiling events (capi.cu): recorded by the launcher between operand preparation
  // and the main kernel when non-null
  cudaEvent_t ev_start = nullptr;
  cudaEvent_t ev_prep_done = nullptr;
  // Resource-preparation pass: do everything a launch does EXCEPT enqueue kernels (set function
  // attributes, which also forces the lazily loaded cubin in).  mm_kernel_execute runs it before
  // recording its start event so that the reported device time is kernel time only.
  bool dry_run = false;
};

// ---- kernel families (one launcher per translation unit) ---------------------------------------
// CUDA-core semiring tile kernel, any (dtype, map, reduce).  semiring_*.cu
int launch_semiring(int dtype, int map_op, int reduce_op, const GemmArgs &args);

// tcgen05 tensor-core GEMM for (Multiply, Add) float (kind::tf32) and half (kind::f16).
// `scratch_a` / `scratch_b` hold the K-major operand copies.  gemm_tcgen05.cu
size_t tcgen05_scratch_bytes(int dtype, unsigned n, unsigned k, unsigned m, int flags);
size_t tcgen05_bt_bytes(int dtype, unsigned k, unsigned m, int flags);  // offset of the A operand copy in scratch
int launch_tcgen05(int dtype, const GemmArgs &args, void *scratch, size_t scratch_bytes);
// The three phases of launch_tcgen05, for callers that reuse a prepared B across row-blocks
// (the pipelined host path, multi-GPU row-block drivers):
// *b_op receives the B operand of tcgen05_gemm: `bt` (prepared K-major copy) or `b` itself when the
// kernel reads the row-major B directly (half).
int tcgen05_prepare_b(int dtype, const void *b, void *bt, unsigned k, unsigned m, int flags,
                      const void **b_op, cudaStream_t stream);
bool tcgen05_b_direct(int dtype);
bool tcgen05_fuse_a(int dtype, int flags);
// *a_raw != nullptr on return: nothing was launched, the GEMM kernel itself rounds *a_raw into *a_op.
int tcgen05_prepare_a(int dtype, const void *a, void *aprep, unsigned rows, unsigned k, int flags,
                      const void **a_op, const void **a_raw, cudaStream_t stream);
// `tile_sync`: device counter (zeroed by the launcher) for the kernel's soft wave barrier, or null.
int tcgen05_gemm(int dtype, const void *a_op, const void *b_op, void *c, unsigned rows, unsigned k,
                 unsigned m, int flags, unsigned int *tile_sync, const void *a_raw, unsigned int *counters,
                 cudaStream_t stream);
constexpr size_t kTcgen05TailBytes = 256 + 64 * 1024;  // [granule counters][wave-barrier counter] at the scratch tail

// DMMA (mma.sync m8n8k4 f64) GEMM for (Multiply, Add) double.  gemm_dmma.cu
int launch_dmma(const GemmArgs &args);

}  // namespace mm
