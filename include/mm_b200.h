/* mm_b200.h — C-ABI of libmm_b200.so: the B200-native MatrixMultiplication hot path.
 *
 * This is the drop-in boundary for ONE path of spcl/gemm_hls: the tiled matrix-multiplication
 * kernel  C[N x M] = A[N x K] (x) B[K x M]  over an (OperatorMap, OperatorReduce) semiring
 * (reference: kernel/Top.cpp:9-117 -> kernel/Memory.cpp streamers -> kernel/Compute.cpp PE chain).
 * Every entry point below names the reference interface it replaces (file:line under the
 * reference checkout).  Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * Layouts (identical to the reference, include/Utility.h:27-39): A row-major N x K (K x N when
 * MM_FLAG_TRANSPOSED_A), B row-major K x M, C row-major N x M, dense, no leading dimensions.
 * The reference's MemoryPack*_t arrays are bit-identical to flat Data_t arrays
 * (include/Utility.h:44-63), so flat pointers are passed here.
 *
 * Error handling: every function returns MM_OK (0) or an MM_ERR_* code; mm_last_error() gives the
 * message for the calling thread.  The library never exits the process and never falls back to a
 * CPU implementation: without a usable CUDA device every compute entry point fails with
 * MM_ERR_CUDA.
 */
#ifndef MM_B200_H_
#define MM_B200_H_

#include <stddef.h>

#if defined(__GNUC__)
#define MM_API __attribute__((visibility("default")))
#else
#define MM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* MM_DATA_TYPE (reference CMakeLists.txt:16, include/Config.h.in:15). */
enum {
  MM_DTYPE_HALF = 0,   /* "half"      IEEE binary16 */
  MM_DTYPE_FLOAT = 1,  /* "float"  */
  MM_DTYPE_DOUBLE = 2, /* "double" */
  MM_DTYPE_INT32 = 3,  /* "int"    */
  MM_DTYPE_UINT32 = 4, /* "unsigned" */
  MM_DTYPE_UINT8 = 5,  /* "uint8_t" (reference CMakeLists.txt:43-46) */
  MM_DTYPE_COUNT = 6
};

/* MM_MAP_OP / MM_REDUCE_OP = hlslib::op functors (hlslib/include/hlslib/xilinx/Operators.h:20-100).
 * identity(): Add 0, Multiply 1, And true, Min numeric_limits<T>::max(),
 * Max numeric_limits<T>::min() (the smallest POSITIVE value for floating point — reproduced). */
enum {
  MM_OP_MULTIPLY = 0, /* hlslib::op::Multiply == Product */
  MM_OP_ADD = 1,      /* hlslib::op::Add == Sum */
  MM_OP_MIN = 2,
  MM_OP_MAX = 3,
  MM_OP_AND = 4,
  MM_OP_COUNT = 5
};

/* Flags for the execute calls. */
enum {
  MM_FLAG_NONE = 0,
  /* A is stored K x N (reference MM_TRANSPOSED_A, CMakeLists.txt:30, include/Utility.h:31-35). */
  MM_FLAG_TRANSPOSED_A = 1,
  /* Force the CUDA-core semiring kernel also for (Multiply, Add), and the literal C++ Min / Max for
   * float.  That kernel accumulates each C element sequentially over k in Data_t with one rounding
   * per Map and per Reduce (no FMA contraction), i.e. it is BIT-IDENTICAL to the reference's Naive<>
   * / FPGA datapath for every data type and every input, including NaN and signed zeros.
   * Without the flag:
   *   - float  (Multiply, Add): tcgen05 kind::tf32, operands rounded to nearest TF32 (10-bit
   *     mantissa), FP32 accumulation.  Meets the reference's 1e-3 relative criterion for SAME-SIGN
   *     data such as its own U[1,10] inputs (measured <= 2.1e-4); for mixed-sign data the error is
   *     bounded relative to sum|a*b|, not to |sum a*b| — use MM_FLAG_TF32X3 or MM_FLAG_EXACT there.
   *   - double (Multiply, Add): DMMA, FP64 throughout; differs from Naive<> only by summation order
   *     (measured <= 1e-14 relative).
   *   - half   (Multiply, Add): tcgen05 kind::f16 with FP32 accumulation and ONE rounding to half at
   *     the end.  The reference (Naive<half>, and its FPGA datapath) accumulates IN HALF, so this
   *     path is closer to the exact product than the reference is and does NOT reproduce the
   *     reference's rounding: against Naive<half> on U[1,10] inputs 15 % of the elements differ by
   *     more than 1e-3 at K = 32 and 63 % at K = 544.  The reference's TestSimulation compares half
   *     EXACTLY (test/TestSimulation.cpp:79-85), so the host executables of this project build half
   *     with MM_FLAG_EXACT unless configured with -DMM_HALF_TENSOR=ON.
   *   - uint8_t (Multiply, Add): tcgen05 kind::i8, exact 32-bit integer accumulation, low byte stored — BIT-IDENTICAL
   *     to the reference's modulo-256 arithmetic.  Used for K <= 33024 (255^2 * K < 2^31); the CUDA-core kernel beyond.
   *   - float Min / Max (as Map or Reduce): the hardware FMNMX.  Identical to the reference's
   *     `(a < b) ? a : b` for all finite inputs except that a tie between -0 and +0 yields -0 for Min
   *     (+0 for Max) where the reference returns the second operand, and NaN operands are dropped
   *     where the reference's comparison lets a NaN in the SECOND operand through.  Inputs without
   *     NaN and without negative zeros (or products / sums that produce them) are bit-identical. */
  MM_FLAG_EXACT = 2,
  /* float (Multiply, Add) on the tensor cores with the 3xTF32 split (hi*hi + hi*lo + lo*hi, each
   * operand split into two TF32 values): ~FP32 accuracy (about 1e-6 relative) at 1/3 of the TF32
   * rate.  Ignored for every other configuration. */
  MM_FLAG_TF32X3 = 4
};

enum {
  MM_OK = 0,
  MM_ERR_INVALID = 1, /* bad argument (null pointer, unknown dtype/op, zero size) */
  MM_ERR_SHAPE = 2,   /* K or M not divisible by the memory width (host/RunHardware.cpp:50-61) */
  MM_ERR_CUDA = 3,    /* CUDA runtime / driver failure, or no device */
  MM_ERR_NOMEM = 4,   /* device allocation failed */
  MM_ERR_UNSUPPORTED = 5
};

/* Opaque device context.  Replaces hlslib::ocl::Context + Program + Kernel
 * (hlslib/include/hlslib/common/OpenCL.h:366-500, host/RunHardware.cpp:116-154): owns the CUDA
 * device binding, one stream, timing events and the scratch the tensor-core path needs. */
typedef struct mm_context mm_context;

/* Message for the last failing call on this thread (never NULL).  Replaces the what() of
 * hlslib::ocl::{ConfigurationError,RuntimeError} (common/OpenCL.h:99-157). */
MM_API const char *mm_last_error(void);

/* Size in bytes of one element of `dtype` (0 if unknown).  Reference: sizeof(Data_t). */
MM_API size_t mm_dtype_size(int dtype);

/* Elements per 64-byte memory word = kMemoryWidthK == kMemoryWidthM
 * (include/MatrixMultiplication.h:18-27 with the default 64-byte bus, CMakeLists.txt:17-19).
 * K and M must be multiples of it (host/RunHardware.cpp:50-61, test/TestSimulation.cpp:22-35). */
MM_API unsigned mm_memory_width(int dtype);

/* hlslib::ocl::Context::Context() (common/OpenCL.h:366-470; host/RunHardware.cpp:116).
 * `device` is the CUDA ordinal (the reference always takes device 0 of the Xilinx platform). */
MM_API int mm_context_create(int device, mm_context **out);
MM_API int mm_context_destroy(mm_context *ctx);

/* Context::MakeBuffer<T, Access>(StorageType, bank, count) (common/OpenCL.h:420-470,502-633;
 * host/RunHardware.cpp:122-138).  Memory banks have no B200 counterpart: one HBM3e pool. */
MM_API int mm_buffer_alloc(mm_context *ctx, size_t bytes, void **device_ptr);
MM_API int mm_buffer_free(mm_context *ctx, void *device_ptr);

/* Buffer::CopyFromHost / Buffer::CopyToHost (common/OpenCL.h:648-720; host/RunHardware.cpp:140-145,
 * 187-190).  Blocking, like the reference. */
MM_API int mm_copy_to_device(mm_context *ctx, void *device_dst, const void *host_src, size_t bytes);
MM_API int mm_copy_to_host(mm_context *ctx, void *host_dst, const void *device_src, size_t bytes);

/* Program::MakeKernel("MatrixMultiplicationKernel", a, b, c, n, k, m) + Kernel::ExecuteTask()
 * (common/OpenCL.h:1346-1379,1486-1504; host/RunHardware.cpp:147-162) on DEVICE buffers.
 * Blocking.  *seconds_device = CUDA-event time around the kernels of this call only (the
 * counterpart of the OpenCL profiling END-START the reference reports); *seconds_wall = host
 * wall clock.  Either may be NULL. */
MM_API int mm_kernel_execute(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                      const void *a_device, const void *b_device, void *c_device, unsigned size_n,
                      unsigned size_k, unsigned size_m, double *seconds_device,
                      double *seconds_wall);

/* Same launch, asynchronous on a caller-supplied CUDA stream (cudaStream_t passed as void*; NULL =
 * the context's own stream), no timing, no synchronisation: for callers that own the stream
 * (benchmark loops, CUDA-graph capture, multi-GPU row-block drivers).  The context's scratch is
 * shared by everything enqueued through it: keep the work of ONE context stream-ordered (one stream
 * at a time) and use one context per concurrently running stream.  Capturable into a CUDA graph
 * once a first call outside capture (or mm_context_reserve) has sized the scratch; see
 * mm_context_reserve for what growth does to captured graphs.  Device pointers must be 16-byte
 * aligned (the reference's buffers are 4096-aligned, include/Utility.h:44-54): MM_ERR_INVALID else. */
MM_API int mm_kernel_enqueue(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                      const void *a_device, const void *b_device, void *c_device, unsigned size_n,
                      unsigned size_k, unsigned size_m, void *cuda_stream);

/* Per-phase device timing of enqueued work, for roofline accounting.  With profiling on, every
 * mm_kernel_enqueue()/mm_kernel_execute() records CUDA events on the launching stream around
 * (i) the operand-preparation kernels and (ii) the main compute kernel.  mm_context_profile_read()
 * synchronises on the recorded events, returns the SUMS over the calls since the last read (at
 * most 256 calls are kept) and the number of calls, and resets the counters. */
MM_API int mm_context_set_profiling(mm_context *ctx, int enable);
MM_API int mm_context_profile_read(mm_context *ctx, double *prep_seconds_sum,
                                   double *main_seconds_sum, int *calls);

/* Number of kernels one mm_kernel_enqueue() with these arguments launches (for accounting). */
MM_API int mm_kernel_launch_count(int dtype, int map_op, int reduce_op, int flags);

/* Name of the compute kernel family mm_kernel_enqueue() would dispatch to for a small problem
 * ("tcgen05_tf32", "tcgen05_f16", "tcgen05_i8", "dmma_f64", "semiring_simt"); static string. */
MM_API const char *mm_kernel_path(int dtype, int map_op, int reduce_op, int flags);

/* The reference's simulation entry, extern "C" MatrixMultiplicationKernel(a, b, c, n, k, m)
 * (include/MatrixMultiplication.h:155-171, called with HOST pointers at
 * test/TestSimulation.cpp:66), for a run-time chosen configuration: host -> device copies,
 * the kernel, device -> host copy of C; blocking.  With `ctx` NULL it uses (and lazily creates) a
 * per-process default: one context on device $MM_DEVICE (default 0), or — when the environment
 * variable MM_NUM_GPUS is G > 1 and A is row-major — an mm_multi over devices 0..G-1, i.e. the
 * call is split over G GPUs exactly like mm_multi_gemm_host().  Timings optional as above;
 * *seconds_device covers the kernels only. */
MM_API int mm_gemm_host(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                 const void *a_host, const void *b_host, void *c_host, unsigned size_n,
                 unsigned size_k, unsigned size_m, double *seconds_device, double *seconds_wall);

/* ---- tuning -------------------------------------------------------------------------------------
 * Run-time counterpart of the reference's configure-time tile / parallelism knobs
 * (CMakeLists.txt:17-29: MM_PARALLELISM_*, MM_MEMORY_TILE_SIZE_*), which scripts/build_manager.py:224-306
 * sweeps by rebuilding.  Here every variant is compiled into the library and selected per context.
 * Defaults are the measured best; the environment variable named with each knob, read ONCE at
 * mm_context_create(), overrides the default; mm_context_set_tuning() overrides both.  Every value
 * computes the same C (the parity suite runs under each). */
enum {
  MM_TUNE_TCGEN05_CTA_GROUP = 0,   /* 1 | 2: single-CTA tiles or cta_group::2 CTA pairs (default 2)    MM_TCGEN05_CTA_GROUP */
  MM_TUNE_TCGEN05_BLOCK_N = 1,     /* 128 | 256: C tile columns = UMMA N (default 256)                 MM_TCGEN05_BLOCK_N */
  MM_TUNE_TCGEN05_STAGES = 2,      /* TMA ring depth, 0 = deepest that fits (default), else 2..8       MM_TCGEN05_STAGES */
  MM_TUNE_TCGEN05_RASTER_ROWS = 3, /* C rows per rasterisation group = L2 "memory tile" height (2048)  MM_TCGEN05_RASTER_ROWS */
  MM_TUNE_TCGEN05_TILE_SYNC = 4,   /* 0 | 1: soft wave barrier between co-running tiles (default 1)    MM_TCGEN05_TILE_SYNC */
  MM_TUNE_TCGEN05_B_MN = 5,        /* 0 | 1: read B MN-major from its row-major layout (1) or K-major
                                      from a transposed copy (0)                                      MM_TCGEN05_B_MN */
  MM_TUNE_TCGEN05_L2_POLICY = 6,   /* TMA loads' L2 eviction priority: 0 normal, 1 first, 2 last       MM_TCGEN05_L2 */
  MM_TUNE_TCGEN05_B_OVERLAP = 7,   /* 0 | 1: prepare B concurrently with the GEMM that consumes it panel by
                                      panel (default 0: measured slower, profiles/r02_exp_b_overlap.md) MM_TCGEN05_B_OVERLAP */
  MM_TUNE_TCGEN05_TMA_STORE = 8,   /* 0 | 1: epilogue through shared memory + TMA stores (default 1)   MM_TCGEN05_TMA_STORE */
  MM_TUNE_DMMA_TILE_ROWS = 9,      /* 0 = automatic | 64 | 128: CTA tile rows of the double kernel     MM_DMMA_TILE_ROWS */
  MM_TUNE_EXPERIMENT_TF32_NO_ROUND = 10, /* 1: feed raw fp32 bits to kind::tf32 (measures the truncation
                                      bias that motivates the rounding pass; never for production)    MM_EXPERIMENT_TF32_NO_ROUND */
  MM_TUNE_SEMIRING_RING = 11,      /* 0 | 1: CUDA-core kernel for 4-byte types with both tiles in a TMA ring and no
                                      block-wide barrier (default 1), or the register-staged kernel       MM_SEMIRING_RING */
  MM_TUNE_COUNT = 12
};
MM_API int mm_context_set_tuning(mm_context *ctx, int knob, int value);
MM_API int mm_context_get_tuning(mm_context *ctx, int knob, int *value);

/* Pre-size the context's scratch for the largest problem that will be enqueued.  The scratch grows
 * on demand otherwise; growth reallocates it, which INVALIDATES any CUDA graph captured through
 * mm_kernel_enqueue() earlier (the old scratch addresses are baked into the graph).  A context that
 * has seen a stream capture therefore keeps superseded allocations alive until it is destroyed. */
MM_API int mm_context_reserve(mm_context *ctx, int dtype, int flags, unsigned size_n, unsigned size_k,
                              unsigned size_m);

/* ---- multi-GPU: C row-blocks over the GPUs of one box ------------------
 * The reference's API is ONE blocking call on host pointers (include/MatrixMultiplication.h:155-171)
 * and one device-resident lifecycle (host/RunHardware.cpp:116-190); outer tiles of C are independent
 * (kernel/Compute.cpp:53-56).  mm_multi keeps both shapes over G devices of this process: GPU g owns
 * rows [g*ceil(N/G), ...) of A and C; every GPU uploads only ITS 1/G row-slice of B over PCIe and the
 * slices are all-gathered GPU-to-GPU over NVLink by the library's own kernels reading peer memory —
 * fused with B's TF32 rounding on the float path, where the GEMM consumes finished panels while the
 * gather is still running.  No collective per step, no reduction (K is not split).  All fan-out is
 * internal (one host thread per GPU) and joined before the call returns.
 * A stored K x N (MM_FLAG_TRANSPOSED_A) cannot be cut into contiguous row blocks: MM_ERR_UNSUPPORTED. */
typedef struct mm_multi mm_multi;
/* `devices` = n_gpus CUDA ordinals, or NULL for 0 .. n_gpus-1. */
MM_API int mm_multi_create(int n_gpus, const int *devices, mm_multi **out);
MM_API int mm_multi_destroy(mm_multi *multi);
MM_API int mm_multi_device_count(const mm_multi *multi);
/* The per-device context (for mm_context_set_tuning); owned by `multi`. */
MM_API mm_context *mm_multi_context(mm_multi *multi, int index);
/* How an N x K x M problem is cut over `n_gpus` devices — pure arithmetic, no device needed; the same rule every
 * mm_multi_* entry applies.  GPU `index` owns rows [*row_begin, *row_end) of A and C (ceil(N / G) rows each, tail GPUs
 * may own fewer or none) and uploads the K-rows [*b_row_begin, *b_row_end) of B (slices of ceil(K / G) rounded up to
 * a multiple of 64 rows; the whole of B when there is no peer access).  Callers use it to place their host matrices
 * (e.g. each block on the NUMA node of the GPU that copies it).  Any output pointer may be NULL. */
MM_API int mm_multi_partition(int n_gpus, int index, unsigned size_n, unsigned size_k, unsigned *row_begin,
                              unsigned *row_end, unsigned *b_row_begin, unsigned *b_row_end);
/* 1 if every pair of devices has peer access (NVLink gather), 0 if B falls back to a full upload per GPU. */
MM_API int mm_multi_peer_access(const mm_multi *multi);
/* MatrixMultiplicationKernel(a, b, c, n, k, m) with HOST pointers over all G devices; blocking.
 * *seconds_device = the slowest GPU's first-kernel-start .. last-kernel-end. */
MM_API int mm_multi_gemm_host(mm_multi *multi, int dtype, int map_op, int reduce_op, int flags,
                              const void *a_host, const void *b_host, void *c_host, unsigned size_n,
                              unsigned size_k, unsigned size_m, double *seconds_device,
                              double *seconds_wall);
/* Device-resident lifecycle, the multi-GPU form of MakeBuffer + CopyFromHost / ExecuteTask / CopyToHost
 * (host/RunHardware.cpp:122-190): upload = A row-blocks + B slices over PCIe, B assembled on every
 * GPU over NVLink; execute = the kernels of every GPU on the resident buffers (device seconds = the
 * slowest GPU's CUDA-event time around its kernels); download = C row-blocks to the host. */
MM_API int mm_multi_upload(mm_multi *multi, int dtype, int flags, const void *a_host, const void *b_host,
                           unsigned size_n, unsigned size_k, unsigned size_m);
MM_API int mm_multi_execute(mm_multi *multi, int dtype, int map_op, int reduce_op, int flags,
                            unsigned size_n, unsigned size_k, unsigned size_m, double *seconds_device,
                            double *seconds_wall);
MM_API int mm_multi_download(mm_multi *multi, int dtype, void *c_host, unsigned size_n, unsigned size_m);

/* Library/ABI version (major * 100 + minor). */
MM_API int mm_version(void);

#ifdef __cplusplus
}
#endif

#endif /* MM_B200_H_ */
