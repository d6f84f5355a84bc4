// C-ABI of libmm_b200.so (include/mm_b200.h): context / buffer lifecycle mirroring the
// hlslib::ocl calls of host/RunHardware.cpp:116-190, and the dispatch of one
// MatrixMultiplicationKernel invocation onto the B200 kernel families.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace mm {

namespace {
thread_local std::string g_last_error = "";
}

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  set_error(msg);
  return code;
}

}  // namespace mm

struct mm_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H streams of the pipelined host path
  std::vector<cudaEvent_t> sync_events;                // untimed events ordering the three streams
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  mm::Scratch scratch;       // operand copies of the tensor-core path
  mm::Scratch staging[3];    // device A, B, C of mm_gemm_host
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // 3 per profiled call: start, after prep, after main
  int prof_calls = 0;
  std::mutex mutex;          // entry points are blocking and serialised per context, like
                             // hlslib::ocl::Context's enqueue/memcopy mutexes (common/OpenCL.h:474-492)
};

namespace {

using mm::fail;

bool valid_dtype(int d) { return d >= 0 && d < MM_DTYPE_COUNT; }
bool valid_op(int o) { return o >= 0 && o < MM_OP_COUNT; }

enum Path { kPathTcgen05, kPathDmma, kPathSemiring };

Path select_path(int dtype, int map_op, int reduce_op, int flags) {
  const bool dense = (map_op == MM_OP_MULTIPLY && reduce_op == MM_OP_ADD) && !(flags & MM_FLAG_EXACT);
  if (dense && (dtype == MM_DTYPE_FLOAT || dtype == MM_DTYPE_HALF)) return kPathTcgen05;
  if (dense && dtype == MM_DTYPE_DOUBLE) return kPathDmma;
  return kPathSemiring;
}

int ensure(mm::Scratch &s, size_t bytes) {
  if (s.bytes >= bytes) return MM_OK;
  if (s.ptr) cudaFree(s.ptr);
  s.ptr = nullptr;
  s.bytes = 0;
  cudaError_t e = cudaMalloc(&s.ptr, bytes);
  if (e != cudaSuccess) {
    return fail(e == cudaErrorMemoryAllocation ? MM_ERR_NOMEM : MM_ERR_CUDA,
                std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  }
  s.bytes = bytes;
  return MM_OK;
}

int check_args(int dtype, int map_op, int reduce_op, const void *a, const void *b, const void *c,
               unsigned n, unsigned k, unsigned m) {
  if (!valid_dtype(dtype)) return fail(MM_ERR_INVALID, "unknown MM_DATA_TYPE code");
  if (!valid_op(map_op) || !valid_op(reduce_op)) return fail(MM_ERR_INVALID, "unknown MM_MAP_OP / MM_REDUCE_OP code");
  if (!a || !b || !c) return fail(MM_ERR_INVALID, "null matrix pointer");
  if (n == 0 || k == 0 || m == 0) return fail(MM_ERR_INVALID, "matrix dimensions must be positive");
  const unsigned w = mm_memory_width(dtype);
  // same rule and wording as host/RunHardware.cpp:50-61
  if (k % w != 0) {
    return fail(MM_ERR_SHAPE, "K (" + std::to_string(k) + ") must be divisable by the memory width in K (" +
                                  std::to_string(w) + ").");
  }
  if (m % w != 0) {
    return fail(MM_ERR_SHAPE, "M (" + std::to_string(m) + ") must be divisable by the memory width in M (" +
                                  std::to_string(w) + ").");
  }
  return MM_OK;
}

int enqueue_locked(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags, const void *a,
                   const void *b, void *c, unsigned n, unsigned k, unsigned m, cudaStream_t stream,
                   bool dry_run = false) {
  mm::GemmArgs g{a, b, c, n, k, m, flags, stream};
  g.dry_run = dry_run;
  cudaEvent_t *pe = nullptr;
  if (!dry_run && ctx->profiling && ctx->prof_calls < 256) {
    while (ctx->prof_events.size() < size_t(3 * (ctx->prof_calls + 1))) {
      cudaEvent_t e;
      MM_CUDA_TRY(cudaEventCreate(&e));
      ctx->prof_events.push_back(e);
    }
    pe = &ctx->prof_events[3 * ctx->prof_calls];
    ++ctx->prof_calls;
    g.ev_start = pe[0];
    g.ev_prep_done = pe[1];
    MM_CUDA_TRY(cudaEventRecord(pe[0], stream));
  }
  int rc_launch = MM_OK;
  Path path = select_path(dtype, map_op, reduce_op, flags);
  if (path == kPathDmma && (flags & MM_FLAG_TRANSPOSED_A) && (n % 2 != 0)) path = kPathSemiring;
  switch (path) {
    case kPathTcgen05: {
      const size_t need = mm::tcgen05_scratch_bytes(dtype, n, k, m, flags);
      int rc = ensure(ctx->scratch, need);
      if (rc != MM_OK) return rc;
      rc_launch = mm::launch_tcgen05(dtype, g, ctx->scratch.ptr, ctx->scratch.bytes);
      break;
    }
    case kPathDmma:
      if (pe) MM_CUDA_TRY(cudaEventRecord(pe[1], stream));
      rc_launch = mm::launch_dmma(g);
      break;
    case kPathSemiring:
      if (pe) MM_CUDA_TRY(cudaEventRecord(pe[1], stream));
      rc_launch = mm::launch_semiring(dtype, map_op, reduce_op, g);
      break;
  }
  if (rc_launch != MM_OK) return rc_launch;
  if (pe) MM_CUDA_TRY(cudaEventRecord(pe[2], stream));
  return MM_OK;
}

std::mutex g_default_mutex;
mm_context *g_default_ctx = nullptr;

}  // namespace

extern "C" {

const char *mm_last_error(void) { return mm::g_last_error.c_str(); }

int mm_version(void) { return 100; }

size_t mm_dtype_size(int dtype) {
  switch (dtype) {
    case MM_DTYPE_HALF: return 2;
    case MM_DTYPE_FLOAT: return 4;
    case MM_DTYPE_DOUBLE: return 8;
    case MM_DTYPE_INT32: return 4;
    case MM_DTYPE_UINT32: return 4;
    case MM_DTYPE_UINT8: return 1;
  }
  return 0;
}

unsigned mm_memory_width(int dtype) {
  const size_t s = mm_dtype_size(dtype);
  return s ? static_cast<unsigned>(64 / s) : 0;
}

int mm_context_create(int device, mm_context **out) {
  if (!out) return fail(MM_ERR_INVALID, "null output pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return fail(MM_ERR_CUDA, std::string("no CUDA device available: ") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (device < 0 || device >= count) return fail(MM_ERR_INVALID, "device ordinal out of range");
  MM_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  MM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return fail(MM_ERR_UNSUPPORTED, std::string("libmm_b200 is built for sm_100a only; device is sm_") +
                                        std::to_string(prop.major) + std::to_string(prop.minor));
  }
  mm_context *ctx = new mm_context();
  ctx->device = device;
  MM_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  MM_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
  MM_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
  MM_CUDA_TRY(cudaEventCreate(&ctx->ev_start));
  MM_CUDA_TRY(cudaEventCreate(&ctx->ev_stop));
  *out = ctx;
  return MM_OK;
}

int mm_context_destroy(mm_context *ctx) {
  if (!ctx) return MM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch.ptr) cudaFree(ctx->scratch.ptr);
  for (auto &s : ctx->staging) {
    if (s.ptr) cudaFree(s.ptr);
  }
  for (auto e : ctx->prof_events) cudaEventDestroy(e);
  for (auto e : ctx->sync_events) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->copy_in);
  cudaStreamDestroy(ctx->copy_out);
  cudaEventDestroy(ctx->ev_start);
  cudaEventDestroy(ctx->ev_stop);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return MM_OK;
}

int mm_buffer_alloc(mm_context *ctx, size_t bytes, void **device_ptr) {
  if (!ctx || !device_ptr) return fail(MM_ERR_INVALID, "null argument");
  *device_ptr = nullptr;
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(device_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    return fail(e == cudaErrorMemoryAllocation ? MM_ERR_NOMEM : MM_ERR_CUDA,
                std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  return MM_OK;
}

int mm_buffer_free(mm_context *ctx, void *device_ptr) {
  if (!ctx) return fail(MM_ERR_INVALID, "null context");
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  MM_CUDA_TRY(cudaFree(device_ptr));
  return MM_OK;
}

int mm_copy_to_device(mm_context *ctx, void *device_dst, const void *host_src, size_t bytes) {
  if (!ctx || !device_dst || !host_src) return fail(MM_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  MM_CUDA_TRY(cudaMemcpyAsync(device_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  MM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM_OK;
}

int mm_copy_to_host(mm_context *ctx, void *host_dst, const void *device_src, size_t bytes) {
  if (!ctx || !host_dst || !device_src) return fail(MM_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  MM_CUDA_TRY(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM_OK;
}

int mm_kernel_enqueue(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                      const void *a, const void *b, void *c, unsigned n, unsigned k, unsigned m,
                      void *cuda_stream) {
  if (!ctx) return fail(MM_ERR_INVALID, "null context");
  int rc = check_args(dtype, map_op, reduce_op, a, b, c, n, k, m);
  if (rc != MM_OK) return rc;
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  return enqueue_locked(ctx, dtype, map_op, reduce_op, flags, a, b, c, n, k, m, s);
}

int mm_kernel_execute(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags,
                      const void *a, const void *b, void *c, unsigned n, unsigned k, unsigned m,
                      double *seconds_device, double *seconds_wall) {
  if (!ctx) return fail(MM_ERR_INVALID, "null context");
  int rc = check_args(dtype, map_op, reduce_op, a, b, c, n, k, m);
  if (rc != MM_OK) return rc;
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  // allocate scratch / load kernels first: the device time below is kernel time only, like the
  // OpenCL profiling interval of the reference's ExecuteTask (common/OpenCL.h:1495-1500)
  rc = enqueue_locked(ctx, dtype, map_op, reduce_op, flags, a, b, c, n, k, m, ctx->stream, /*dry_run=*/true);
  if (rc != MM_OK) return rc;
  const auto t0 = std::chrono::high_resolution_clock::now();
  MM_CUDA_TRY(cudaEventRecord(ctx->ev_start, ctx->stream));
  rc = enqueue_locked(ctx, dtype, map_op, reduce_op, flags, a, b, c, n, k, m, ctx->stream);
  if (rc != MM_OK) return rc;
  MM_CUDA_TRY(cudaEventRecord(ctx->ev_stop, ctx->stream));
  MM_CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
  const auto t1 = std::chrono::high_resolution_clock::now();
  float ms = 0.f;
  MM_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
  if (seconds_device) *seconds_device = 1e-3 * ms;
  if (seconds_wall) *seconds_wall = std::chrono::duration<double>(t1 - t0).count();
  return MM_OK;
}

int mm_context_set_profiling(mm_context *ctx, int enable) {
  if (!ctx) return fail(MM_ERR_INVALID, "null context");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  ctx->profiling = enable != 0;
  ctx->prof_calls = 0;
  return MM_OK;
}

int mm_context_profile_read(mm_context *ctx, double *prep_sum, double *main_sum, int *calls) {
  if (!ctx) return fail(MM_ERR_INVALID, "null context");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  double prep = 0.0, main_s = 0.0;
  for (int i = 0; i < ctx->prof_calls; ++i) {
    cudaEvent_t *e = &ctx->prof_events[3 * i];
    MM_CUDA_TRY(cudaEventSynchronize(e[2]));
    float ms_prep = 0.f, ms_main = 0.f;
    MM_CUDA_TRY(cudaEventElapsedTime(&ms_prep, e[0], e[1]));
    MM_CUDA_TRY(cudaEventElapsedTime(&ms_main, e[1], e[2]));
    prep += 1e-3 * ms_prep;
    main_s += 1e-3 * ms_main;
  }
  if (prep_sum) *prep_sum = prep;
  if (main_sum) *main_sum = main_s;
  if (calls) *calls = ctx->prof_calls;
  ctx->prof_calls = 0;
  return MM_OK;
}

int mm_kernel_launch_count(int dtype, int map_op, int reduce_op, int flags) {
  if (!valid_dtype(dtype) || !valid_op(map_op) || !valid_op(reduce_op)) return -1;
  switch (select_path(dtype, map_op, reduce_op, flags)) {
    case kPathTcgen05:
      // [B^T prep unless B is read directly] + [A prep for float or transposed A] + GEMM
      return 1 + (mm::tcgen05_b_direct(dtype) ? 0 : 1) +
             ((dtype == MM_DTYPE_FLOAT || (flags & MM_FLAG_TRANSPOSED_A)) ? 1 : 0);
    case kPathDmma: return 1;
    case kPathSemiring: return 1;
  }
  return -1;
}

const char *mm_kernel_path(int dtype, int map_op, int reduce_op, int flags) {
  if (!valid_dtype(dtype) || !valid_op(map_op) || !valid_op(reduce_op)) return "invalid";
  switch (select_path(dtype, map_op, reduce_op, flags)) {
    case kPathTcgen05: return dtype == MM_DTYPE_FLOAT ? "tcgen05_tf32" : "tcgen05_f16";
    case kPathDmma: return "dmma_f64";
    case kPathSemiring: return "semiring_simt";
  }
  return "invalid";
}

int mm_gemm_host(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags, const void *a,
                 const void *b, void *c, unsigned n, unsigned k, unsigned m, double *seconds_device,
                 double *seconds_wall) {
  int rc = check_args(dtype, map_op, reduce_op, a, b, c, n, k, m);
  if (rc != MM_OK) return rc;
  if (!ctx) {
    std::lock_guard<std::mutex> lock(g_default_mutex);
    if (!g_default_ctx) {
      rc = mm_context_create(0, &g_default_ctx);
      if (rc != MM_OK) return rc;
    }
    ctx = g_default_ctx;
  }
  const auto t0 = std::chrono::high_resolution_clock::now();
  std::lock_guard<std::mutex> lock(ctx->mutex);
  MM_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t es = mm_dtype_size(dtype);
  const size_t bytes_a = size_t(n) * k * es, bytes_b = size_t(k) * m * es, bytes_c = size_t(n) * m * es;
  if ((rc = ensure(ctx->staging[0], bytes_a)) != MM_OK) return rc;
  if ((rc = ensure(ctx->staging[1], bytes_b)) != MM_OK) return rc;
  if ((rc = ensure(ctx->staging[2], bytes_c)) != MM_OK) return rc;
  unsigned char *da = static_cast<unsigned char *>(ctx->staging[0].ptr);
  unsigned char *db = static_cast<unsigned char *>(ctx->staging[1].ptr);
  unsigned char *dc = static_cast<unsigned char *>(ctx->staging[2].ptr);
  const unsigned char *ha = static_cast<const unsigned char *>(a);
  unsigned char *hc = static_cast<unsigned char *>(c);

  // Row-chunk pipeline: C row-blocks are independent (kernel/Compute.cpp:53-56), so the H2D copy of
  // A chunk i+1, the kernels of chunk i and the D2H copy of C chunk i-1 run concurrently on three
  // streams; B is copied (and, on the tcgen05 path, prepared) once up front.  A stored K x N cannot
  // be cut into contiguous row chunks: it takes the single-chunk route.
  Path path = select_path(dtype, map_op, reduce_op, flags);
  const bool ta = (flags & MM_FLAG_TRANSPOSED_A) != 0;
  if (path == kPathDmma && ta && (n % 2 != 0)) path = kPathSemiring;
  unsigned chunk_rows = n;
  if (!ta) {
    const size_t row_bytes = size_t(k) * es;
    size_t rows = std::max<size_t>((n + 15) / 16, ((size_t(32) << 20) + row_bytes - 1) / row_bytes);
    rows = (rows + 127) / 128 * 128;
    if (rows < n) chunk_rows = unsigned(rows);
  }
  const unsigned chunks = (n + chunk_rows - 1) / chunk_rows;
  while (ctx->sync_events.size() < size_t(2 * chunks + 1)) {
    cudaEvent_t e;
    MM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->sync_events.push_back(e);
  }
  cudaEvent_t ev_b = ctx->sync_events[0];
  auto ev_a = [&](unsigned i) { return ctx->sync_events[1 + i]; };
  auto ev_c = [&](unsigned i) { return ctx->sync_events[1 + chunks + i]; };

  // ---- H2D stream
  MM_CUDA_TRY(cudaMemcpyAsync(db, b, bytes_b, cudaMemcpyHostToDevice, ctx->copy_in));
  MM_CUDA_TRY(cudaEventRecord(ev_b, ctx->copy_in));
  for (unsigned i = 0; i < chunks; ++i) {
    const size_t r0 = size_t(i) * chunk_rows, rows = std::min<size_t>(chunk_rows, n - r0);
    if (!ta) {
      MM_CUDA_TRY(cudaMemcpyAsync(da + r0 * k * es, ha + r0 * k * es, rows * k * es, cudaMemcpyHostToDevice,
                                  ctx->copy_in));
    } else {
      MM_CUDA_TRY(cudaMemcpyAsync(da, ha, bytes_a, cudaMemcpyHostToDevice, ctx->copy_in));
    }
    MM_CUDA_TRY(cudaEventRecord(ev_a(i), ctx->copy_in));
  }

  // ---- compute stream
  void *bt = nullptr;
  unsigned char *aprep = nullptr;
  if (path == kPathTcgen05) {
    if ((rc = ensure(ctx->scratch, mm::tcgen05_scratch_bytes(dtype, n, k, m, flags))) != MM_OK) return rc;
    bt = ctx->scratch.ptr;
    aprep = static_cast<unsigned char *>(ctx->scratch.ptr) + mm::tcgen05_bt_bytes(dtype, k, m, flags);
  }
  MM_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev_b, 0));
  MM_CUDA_TRY(cudaEventRecord(ctx->ev_start, ctx->stream));
  const void *b_op = nullptr;
  if (path == kPathTcgen05) {
    if ((rc = mm::tcgen05_prepare_b(dtype, db, bt, k, m, flags, &b_op, ctx->stream)) != MM_OK) return rc;
  }
  for (unsigned i = 0; i < chunks; ++i) {
    const size_t r0 = size_t(i) * chunk_rows, rows = std::min<size_t>(chunk_rows, n - r0);
    MM_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev_a(i), 0));
    const void *a_chunk = da + (ta ? 0 : r0 * k * es);
    void *c_chunk = dc + r0 * m * es;
    if (path == kPathTcgen05) {
      const void *a_op = nullptr, *a_raw = nullptr;
      const size_t a_scale = (dtype == MM_DTYPE_FLOAT && (flags & MM_FLAG_TF32X3)) ? 3 : 1;
      rc = mm::tcgen05_prepare_a(dtype, a_chunk, aprep + (ta ? 0 : r0 * k * es * a_scale), unsigned(rows), k, flags,
                                 &a_op, &a_raw, ctx->stream);
      if (rc != MM_OK) return rc;
      unsigned char *tail = static_cast<unsigned char *>(ctx->scratch.ptr) + ctx->scratch.bytes;
      unsigned int *tile_sync = reinterpret_cast<unsigned int *>(tail - 256);
      unsigned int *a_done = reinterpret_cast<unsigned int *>(tail - mm::kTcgen05TailBytes);
      rc = mm::tcgen05_gemm(dtype, a_op, b_op, c_chunk, unsigned(rows), k, m, flags, tile_sync, a_raw, a_done,
                            ctx->stream);
    } else {
      mm::GemmArgs g{a_chunk, db, c_chunk, unsigned(rows), k, m, flags, ctx->stream};
      rc = (path == kPathDmma) ? mm::launch_dmma(g) : mm::launch_semiring(dtype, map_op, reduce_op, g);
    }
    if (rc != MM_OK) return rc;
    MM_CUDA_TRY(cudaEventRecord(ev_c(i), ctx->stream));
  }
  MM_CUDA_TRY(cudaEventRecord(ctx->ev_stop, ctx->stream));

  // ---- D2H stream
  for (unsigned i = 0; i < chunks; ++i) {
    const size_t r0 = size_t(i) * chunk_rows, rows = std::min<size_t>(chunk_rows, n - r0);
    MM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ev_c(i), 0));
    MM_CUDA_TRY(cudaMemcpyAsync(hc + r0 * m * es, dc + r0 * m * es, rows * m * es, cudaMemcpyDeviceToHost,
                                ctx->copy_out));
  }
  MM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_out));
  MM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (seconds_device) {
    // first kernel start .. last kernel end on the compute stream (with more than one chunk this
    // includes the stalls waiting for A chunks to arrive)
    float ms = 0.f;
    MM_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
    *seconds_device = 1e-3 * ms;
  }
  if (seconds_wall) {
    *seconds_wall = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  }
  return MM_OK;
}

}  // extern "C"
