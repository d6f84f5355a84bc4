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

// Run-time tuning of the kernel families: the B200 counterpart of the reference's CMake-time tile /
// parallelism knobs (CMakeLists.txt:17-29), swept by scripts/tile_sweep.py the way
// scripts/build_manager.py:224-306 sweeps builds.  One instance per context: defaults, then the
// MM_TUNE_* environment variables read ONCE at mm_context_create(), then mm_context_set_tuning().
// Indexed by the MM_TUNE_* codes of include/mm_b200.h.
struct Tuning {
  int v[MM_TUNE_COUNT];
  int cta_group() const { return v[MM_TUNE_TCGEN05_CTA_GROUP]; }
  int block_n() const { return v[MM_TUNE_TCGEN05_BLOCK_N]; }
  int stages() const { return v[MM_TUNE_TCGEN05_STAGES]; }
  int raster_rows() const { return v[MM_TUNE_TCGEN05_RASTER_ROWS]; }
  bool tile_sync() const { return v[MM_TUNE_TCGEN05_TILE_SYNC] != 0; }
  bool b_mn() const { return v[MM_TUNE_TCGEN05_B_MN] != 0; }
  int l2_policy() const { return v[MM_TUNE_TCGEN05_L2_POLICY]; }
  int b_overlap() const { return v[MM_TUNE_TCGEN05_B_OVERLAP]; }
  bool tma_store() const { return v[MM_TUNE_TCGEN05_TMA_STORE] != 0; }
  int dmma_tile_rows() const { return v[MM_TUNE_DMMA_TILE_ROWS]; }
  bool tf32_no_round() const { return v[MM_TUNE_EXPERIMENT_TF32_NO_ROUND] != 0; }
  bool semiring_ring() const { return v[MM_TUNE_SEMIRING_RING] != 0; }
};
Tuning default_tuning();                                  // capi.cu
int tuning_validate(int knob, int value);                 // MM_OK or MM_ERR_INVALID (message set)

struct GemmArgs {
  const void *a;
  const void *b;
  void *c;
  unsigned n, k, m;
  int flags;
  cudaStream_t stream;
  const Tuning *tuning = nullptr;  // never null on a real launch (capi.cu fills it from the context)
  // Second stream + fork/join events of the context: B's operand preparation runs there, overlapped
  // with the GEMM that consumes it panel by panel (gemm_tcgen05.cu).  Null = no overlap.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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
// The context's scratch holds, in this order: [B operand copy][A operand copy][counters].
//   B operand copy: float = B rounded to TF32 (same row-major K x M layout when the kernel reads B
//     MN-major, the transposed M x K copy otherwise); half = nothing when B is read in place.
//   A operand copy: float = A rounded to TF32; any type with MM_FLAG_TRANSPOSED_A = A transposed.
size_t tcgen05_scratch_bytes(int dtype, unsigned n, unsigned k, unsigned m, int flags, const Tuning &t);
size_t tcgen05_bt_bytes(int dtype, unsigned k, unsigned m, int flags, const Tuning &t);  // = offset of the A copy
int launch_tcgen05(int dtype, const GemmArgs &args, void *scratch, size_t scratch_bytes);
// How the kernel consumes B for this configuration.
bool tcgen05_b_mn(int dtype, int flags, const Tuning &t);      // MN-major (row-major K x M array) vs K-major copy
bool tcgen05_b_in_place(int dtype, int flags, const Tuning &t);  // no B copy at all (half, MN-major)

// The phases of launch_tcgen05, for callers that reuse a prepared B across row-blocks (the pipelined
// host path, the multi-GPU row-block driver).
// Where the kernel's B operand comes from.  `src` non-null: `parts` row-slices of B of `part_rows`
// rows each (the last may be short), slice j read through src[j] — a pointer to a FULL-size K x M
// array of which only slice j's rows need to be valid (peer-GPU buffers of the multi-GPU path: the
// NVLink all-gather of B is fused into this pass).  `src` null: one array, `b`.
struct BSource {
  const void *b = nullptr;
  const void *const *src = nullptr;  // DEVICE array of `parts` pointers
  unsigned parts = 1;
  unsigned part_rows = 0;
};
// Counters at the tail of the scratch (zeroed by the launchers before use).
struct Tcgen05Counters {
  unsigned int *tile_sync;  // soft wave barrier of the GEMM
  unsigned int *b_ready;    // one per BLOCK_N-column panel of B: finished preparation work items
};
Tcgen05Counters tcgen05_counters(void *scratch, size_t scratch_bytes);
// Prepares B into `bt` on `stream`.  *b_op receives the kernel's B operand (bt, or the caller's B when
// it is read in place).  With `ready` non-null the pass runs panel by panel (BLOCK_N columns of B at
// a time, in the order the GEMM's rasterisation consumes them) as a co-resident persistent kernel and
// publishes each panel through ready[panel]; *ready_target receives the count that means "complete".
int tcgen05_prepare_b(int dtype, const BSource &src, void *bt, unsigned k, unsigned m, int flags, const Tuning &t,
                      const void **b_op, unsigned int *ready, unsigned *ready_target, cudaStream_t stream);
// Fork / join wrapper around tcgen05_prepare_b for the launchers: decides whether the preparation
// overlaps the GEMM (float rounding, or any gather of peer slices, with the MN-major B path and a side
// stream), zeroes the panel counters in stream order, runs the pass on `side` — ENQUEUED BEFORE the
// GEMM that waits on its counters, so that serialising tools (ncu, compute-sanitizer) still run it
// first — and records `ev_join` there.  `local_b`: where a plain gather of slices goes when B needs no
// scratch copy (half); may be null for a single source.
struct PreparedB {
  const void *b_op = nullptr;
  const unsigned int *ready = nullptr;  // pass to tcgen05_gemm
  unsigned ready_target = 0;
  bool forked = false;  // the caller makes `stream` wait on `ev_join` after its last GEMM
};
int tcgen05_prepare_b_async(int dtype, const BSource &src, void *local_b, void *scratch, size_t scratch_bytes,
                            unsigned k, unsigned m, int flags, const Tuning &t, cudaStream_t stream, cudaStream_t side,
                            cudaEvent_t ev_fork, cudaEvent_t ev_join, PreparedB *out);
int tcgen05_prepare_a(int dtype, const void *a, void *aprep, unsigned rows, unsigned k, int flags, const Tuning &t,
                      const void **a_op, cudaStream_t stream);
// `tile_sync`: device counter for the kernel's soft wave barrier, or null.  `b_ready` non-null: the
// producer waits for b_ready[column tile] >= b_ready_target before it fetches a tile's B panel.
int tcgen05_gemm(int dtype, const void *a_op, const void *b_op, void *c, unsigned rows, unsigned k, unsigned m,
                 int flags, const Tuning &t, unsigned int *tile_sync, const unsigned int *b_ready,
                 unsigned b_ready_target, cudaStream_t stream);
constexpr size_t kTcgen05TailBytes = 256 + 64 * 1024;  // [panel counters, 64 KiB][wave-barrier counter, 256 B]
// Generic gather of row-sliced B into one local array (identity transform): what the multi-GPU path
// uses for the kernel families that read B as is (double, semirings, half).
int gather_b_rows(const BSource &src, void *dst, size_t elem_bytes, unsigned k, unsigned m, cudaStream_t stream);

// DMMA (mma.sync m8n8k4 f64) GEMM for (Multiply, Add) double.  gemm_dmma.cu
int launch_dmma(const GemmArgs &args);

}  // namespace mm
