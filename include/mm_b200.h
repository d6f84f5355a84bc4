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
  /* Force the CUDA-core semiring kernel also for (Multiply, Add).  That kernel accumulates each
   * C element sequentially over k in Data_t with one rounding per Map and per Reduce (no FMA
   * contraction), i.e. it is BIT-IDENTICAL to the reference's Naive<> / FPGA datapath for every
   * data type.  Without the flag, half/float/double (Multiply, Add) use the tensor cores
   * (tcgen05 kind::f16 / kind::tf32 with FP32 accumulation in TMEM; DMMA for double), which is
   * within the reference's 1e-3 relative tolerance but not bit-identical. */
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
 * once a first call outside capture has sized the scratch. */
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

/* Name of the compute kernel family mm_kernel_enqueue() would dispatch to
 * ("tcgen05_tf32", "tcgen05_f16", "dmma_f64", "semiring_simt"); static string. */
MM_API const char *mm_kernel_path(int dtype, int map_op, int reduce_op, int flags);

/* The reference's simulation entry, extern "C" MatrixMultiplicationKernel(a, b, c, n, k, m)
 * (include/MatrixMultiplication.h:155-171, called with HOST pointers at
 * test/TestSimulation.cpp:66), for a run-time chosen configuration: host -> device copies,
 * the kernel, device -> host copy of C; blocking.  Uses (and lazily creates) a per-process
 * default context on device 0, or `ctx` when non-NULL.  Timings optional as above;
 * *seconds_device covers the kernels only. */
MM_API int mm_gemm_host(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                 const void *a_host, const void *b_host, void *c_host, unsigned size_n,
                 unsigned size_k, unsigned size_m, double *seconds_device, double *seconds_wall);

/* Library/ABI version (major * 100 + minor). */
MM_API int mm_version(void);

#ifdef __cplusplus
}
#endif

#endif /* MM_B200_H_ */
