// Tensor-core path of the hot path for the dense (Multiply, Add) contraction on float and half:
//   C[N x M] = A[N x K] * B[K x M]
// B200 counterpart of the reference's PE chain + streamers (kernel/Compute.cpp:53-146,
// kernel/Memory.cpp:58-438) for MM_MAP_OP=Multiply, MM_REDUCE_OP=Add, MM_DATA_TYPE in {float, half}.
//
// Structure (one persistent CTA per SM, warp-specialised, no CUTLASS):
//   warp 0   TMA producer: cp.async.bulk.tensor 128-byte-swizzled A (128 x BK) and B^T (256 x BK)
//            tiles into a STAGES-deep shared-memory ring, mbarrier full/empty pairs
//   warp 1   MMA issuer: one thread issues tcgen05.mma (kind::tf32 | kind::f16, M=128, N=256,
//            K=32 bytes) into one of two 256-column FP32 accumulators in TMEM; tcgen05.commit
//            releases smem stages and publishes finished accumulators
//   warps 2-5 epilogue: tcgen05.ld the accumulator (each warp its 32-lane TMEM quarter), convert,
//            predicated 128-bit stores of the C tile (n < N, m < M masking = WriteC,
//            kernel/Memory.cpp:378-381); overlaps the next tile's main loop (double-buffered TMEM)
//
// Operand preparation (prep kernels below; O(N*K + K*M) bytes against O(N*K*M) flops):
//   * kind::tf32 reads only the upper 19 bits of each fp32 operand, i.e. it TRUNCATES.  The
//     reference's inputs are all positive (U[1,10], test/TestSimulation.cpp:46-55), so truncation
//     would bias every product by about -2^-11 * 2 and land the sum right at the 1e-3 tolerance.
//     A and B are therefore rounded to nearest TF32 (cvt.rna.tf32.f32) first.
//   * both MMA operands are consumed K-major, so B (row-major K x M) is transposed to M x K in the
//     same pass that rounds it (for half: transposed only); A (row-major N x K) is K-major already.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace mm {
namespace {

constexpr int BLOCK_M = 128;          // C rows per tile  (UMMA M)
constexpr int BLOCK_N = 256;          // C cols per tile  (UMMA N)
constexpr int BLOCK_K_BYTES = 128;    // one 128-byte swizzle atom of K per stage
constexpr int UMMA_K_BYTES = 32;      // K extent of one tcgen05.mma
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K_BYTES;  // 16 KiB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K_BYTES;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int RASTER_GROUP = 16;  // row-tiles per rasterisation group (L2 reuse of B^T tiles)
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TileCoord {
  uint32_t r, c;
};

// Grouped rasterisation: RASTER_GROUP row-tiles sweep all column-tiles together so that the ~148
// concurrently running tiles share A row-panels and B column-panels through L2.
__device__ __forceinline__ TileCoord tile_coord(uint32_t t, uint32_t tiles_r, uint32_t tiles_c) {
  const uint32_t per_group = RASTER_GROUP * tiles_c;
  const uint32_t g = t / per_group;
  const uint32_t first = g * RASTER_GROUP;
  const uint32_t gsize = min(static_cast<uint32_t>(RASTER_GROUP), tiles_r - first);
  const uint32_t in = t - g * per_group;
  return TileCoord{first + in % gsize, in / gsize};
}

template <typename TOut>
__device__ __forceinline__ void store_chunk(TOut *crow, const uint32_t (&v)[32], uint32_t col,
                                            uint32_t cols);

template <>
__device__ __forceinline__ void store_chunk<float>(float *crow, const uint32_t (&v)[32],
                                                   uint32_t col, uint32_t cols) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if (col + j + 4 <= cols) {
      *reinterpret_cast<uint4 *>(crow + col + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}

template <>
__device__ __forceinline__ void store_chunk<__half>(__half *crow, const uint32_t (&v)[32],
                                                    uint32_t col, uint32_t cols) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    if (col + j + 8 <= cols) {
      uint32_t p[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __half2 h = __floats2half2_rn(__uint_as_float(v[j + 2 * q]), __uint_as_float(v[j + 2 * q + 1]));
        p[q] = *reinterpret_cast<uint32_t *>(&h);
      }
      *reinterpret_cast<uint4 *>(crow + col + j) = make_uint4(p[0], p[1], p[2], p[3]);
    }
  }
}

// C[rows x cols] = A'[rows x k] * Bt[cols x k]^T ; A', Bt K-major, described by the tensor maps.
template <int KIND, typename TOut>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, TOut *__restrict__ C, uint32_t rows,
                    uint32_t cols, uint32_t k_bytes) {
  extern __shared__ unsigned char smem_raw[];
  // 128B-swizzled tiles must start on a 1024-byte boundary.
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  // generic pointer to the TMEM base slot
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;

  const uint32_t tiles_r = (rows + BLOCK_M - 1) / BLOCK_M;
  const uint32_t tiles_c = (cols + BLOCK_N - 1) / BLOCK_N;
  const uint32_t num_tiles = tiles_r * tiles_c;
  const uint32_t num_kb = (k_bytes + BLOCK_K_BYTES - 1) / BLOCK_K_BYTES;
  constexpr int ELEM_BYTES = (KIND == ptx::KIND_TF32) ? 4 : 2;
  constexpr int BLOCK_K_ELEMS = BLOCK_K_BYTES / ELEM_BYTES;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);   // producer's arrive.expect_tx
      ptx::mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);   // tcgen05.commit
      ptx::mbar_init(tmem_empty_bar(s), 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<1>(tmem_slot, TMEM_COLS);
  }
  ptx::tcgen05_fence_before_sync();
  __syncthreads();
  ptx::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(t, tiles_r, tiles_c);
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          ptx::mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          ptx::tma_load_2d(smem_a0 + stage * A_STAGE_BYTES, &tmap_a, full_bar(stage),
                           kb * BLOCK_K_ELEMS, tc.r * BLOCK_M);
          ptx::tma_load_2d(smem_b0 + stage * B_STAGE_BYTES, &tmap_b, full_bar(stage),
                           kb * BLOCK_K_ELEMS, tc.c * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(KIND, BLOCK_M, BLOCK_N);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++iter) {
        const uint32_t as = iter & 1u;
        const uint32_t aphase = (iter >> 1) & 1u;
        ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        ptx::tcgen05_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after_sync();
          const uint64_t adesc = ptx::make_smem_desc_k_sw128(smem_a0 + stage * A_STAGE_BYTES);
          const uint64_t bdesc = ptx::make_smem_desc_k_sw128(smem_b0 + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K_BYTES / UMMA_K_BYTES; ++k) {
            // advancing K inside the swizzle atom = advancing the start address (>>4 units)
            ptx::umma<KIND, 1>(tmem_d, adesc + uint64_t(k * (UMMA_K_BYTES >> 4)),
                               bdesc + uint64_t(k * (UMMA_K_BYTES >> 4)), idesc,
                               (kb | uint32_t(k)) != 0u ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
          if (kb == num_kb - 1) ptx::umma_commit(tmem_full_bar(as));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const uint32_t quarter = warp & 3u;  // TMEM lanes [32*quarter, +32) are this warp's
    uint32_t iter = 0;
    for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++iter) {
      const TileCoord tc = tile_coord(t, tiles_r, tiles_c);
      const uint32_t as = iter & 1u;
      const uint32_t aphase = (iter >> 1) & 1u;
      ptx::mbar_wait(tmem_full_bar(as), aphase);
      ptx::tcgen05_fence_after_sync();
      const uint32_t row = tc.r * BLOCK_M + quarter * 32 + lane;
      TOut *crow = C + size_t(row) * cols;
      const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + as * BLOCK_N;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr0 + chunk * 32, v);
        ptx::tmem_ld_wait();
        if (row < rows) store_chunk<TOut>(crow, v, tc.c * BLOCK_N + chunk * 32, cols);
      }
      ptx::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(as));
    }
  }

  ptx::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after_sync();
    ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

// ---- operand preparation ------------------------------------------------------------------------
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// dst[i] = rna_tf32(src[i]); count is a multiple of 4 (K % 16 == 0).
__global__ void __launch_bounds__(256)
round_tf32_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t count4) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count4; i += stride) {
    float4 v = src[i];
    v.x = round_tf32(v.x);
    v.y = round_tf32(v.y);
    v.z = round_tf32(v.z);
    v.w = round_tf32(v.w);
    dst[i] = v;
  }
}

template <typename T, bool ROUND>
__device__ __forceinline__ T prep_value(T x) {
  return x;
}
template <>
__device__ __forceinline__ float prep_value<float, true>(float x) {
  return round_tf32(x);
}

// dst[c][r] = f(src[r][c]) for src of shape src_rows x src_cols (row-major): 64 x 64 tiles through
// shared memory so that both the reads and the writes are row-contiguous.
template <typename T, bool ROUND>
__global__ void __launch_bounds__(256)
transpose_prep_kernel(const T *__restrict__ src, T *__restrict__ dst, uint32_t src_rows,
                      uint32_t src_cols) {
  constexpr int TILE = 64;
  constexpr int PAD = (sizeof(T) >= 4) ? 1 : 2;
  __shared__ T tile[TILE][TILE + PAD];
  const uint32_t c0 = blockIdx.x * TILE;
  const uint32_t r0 = blockIdx.y * TILE;
  const int x = threadIdx.x % TILE;
  const int y = threadIdx.x / TILE;  // 0..3
#pragma unroll 4
  for (int i = y; i < TILE; i += 4) {
    const uint32_t r = r0 + i, c = c0 + x;
    if (r < src_rows && c < src_cols) tile[i][x] = prep_value<T, ROUND>(src[size_t(r) * src_cols + c]);
  }
  __syncthreads();
#pragma unroll 4
  for (int i = y; i < TILE; i += 4) {
    const uint32_t c = c0 + i, r = r0 + x;  // dst row = src col
    if (c < src_cols && r < src_rows) dst[size_t(c) * src_rows + r] = tile[x][i];
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// K-major operand: `rows` rows of `k_elems` elements, row pitch k_elems * elem_bytes;
// box = {128 bytes of K, box_rows}, 128-byte swizzle, out-of-bounds reads return zeros
// (neutral for (Multiply, Add) — SURVEY.md section 5 trap 3).
int make_operand_map(CUtensorMap *map, const void *base, int dtype, uint64_t rows, uint64_t k_elems,
                     uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint32_t eb = (dtype == MM_DTYPE_FLOAT) ? 4 : 2;
  cuuint64_t gdim[2] = {k_elems, rows};
  cuuint64_t gstride[1] = {k_elems * eb};
  cuuint32_t box[2] = {uint32_t(BLOCK_K_BYTES / eb), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == MM_DTYPE_FLOAT ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  }
  return MM_OK;
}

int num_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T, bool ROUND>
void launch_transpose(const void *src, void *dst, uint32_t src_rows, uint32_t src_cols,
                      cudaStream_t stream) {
  dim3 grid((src_cols + 63) / 64, (src_rows + 63) / 64);
  transpose_prep_kernel<T, ROUND><<<grid, 256, 0, stream>>>(static_cast<const T *>(src),
                                                           static_cast<T *>(dst), src_rows, src_cols);
}

}  // namespace

size_t tcgen05_scratch_bytes(int dtype, unsigned n, unsigned k, unsigned m, int flags) {
  const size_t eb = (dtype == MM_DTYPE_FLOAT) ? 4 : 2;
  size_t bytes = align_up(size_t(m) * k * eb, 1024);  // B^T
  if (dtype == MM_DTYPE_FLOAT || (flags & MM_FLAG_TRANSPOSED_A)) bytes += align_up(size_t(n) * k * eb, 1024);
  return bytes;
}

int launch_tcgen05(int dtype, const GemmArgs &g, void *scratch, size_t scratch_bytes) {
  if (dtype != MM_DTYPE_FLOAT && dtype != MM_DTYPE_HALF) {
    return fail(MM_ERR_UNSUPPORTED, "tcgen05 path handles float and half only");
  }
  if (g.flags & MM_FLAG_TF32X3) return fail(MM_ERR_UNSUPPORTED, "MM_FLAG_TF32X3 is not implemented yet");
  const bool is_f32 = dtype == MM_DTYPE_FLOAT;
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  const size_t eb = is_f32 ? 4 : 2;
  if (scratch_bytes < tcgen05_scratch_bytes(dtype, g.n, g.k, g.m, g.flags)) {
    return fail(MM_ERR_INVALID, "tcgen05 scratch too small");
  }
  unsigned char *sp = static_cast<unsigned char *>(scratch);
  void *bt = sp;
  void *aprep = sp + align_up(size_t(g.m) * g.k * eb, 1024);

  // ---- operand preparation ----
  const void *a_op = g.a;
  // Experiment hook (scripts/exp_tf32_rounding.py): feed raw fp32 bits to kind::tf32 to MEASURE the
  // truncation bias that motivates the rounding pass.  Never set in production.
  static const bool no_round = std::getenv("MM_EXPERIMENT_TF32_NO_ROUND") != nullptr;
  if (is_f32 && no_round) {
    launch_transpose<float, false>(g.b, bt, g.k, g.m, g.stream);
    if (ta) {
      launch_transpose<float, false>(g.a, aprep, g.k, g.n, g.stream);
      a_op = aprep;
    }
  } else if (is_f32) {
    launch_transpose<float, true>(g.b, bt, g.k, g.m, g.stream);  // B (K x M) -> B^T (M x K), rounded
    if (ta) {
      launch_transpose<float, true>(g.a, aprep, g.k, g.n, g.stream);  // A stored K x N -> N x K
    } else {
      const size_t count4 = size_t(g.n) * g.k / 4;
      const int blocks = int(std::min<size_t>((count4 + 255) / 256, size_t(num_sms()) * 16));
      round_tf32_kernel<<<blocks, 256, 0, g.stream>>>(static_cast<const float4 *>(g.a),
                                                     static_cast<float4 *>(aprep), count4);
    }
    a_op = aprep;
  } else {
    launch_transpose<__half, false>(g.b, bt, g.k, g.m, g.stream);
    if (ta) {
      launch_transpose<__half, false>(g.a, aprep, g.k, g.n, g.stream);
      a_op = aprep;
    }
  }
  MM_CUDA_TRY(cudaGetLastError());
  if (g.ev_prep_done) MM_CUDA_TRY(cudaEventRecord(g.ev_prep_done, g.stream));

  // ---- tensor maps + GEMM ----
  CUtensorMap map_a, map_b;
  int rc = make_operand_map(&map_a, a_op, dtype, g.n, g.k, BLOCK_M);
  if (rc != MM_OK) return rc;
  rc = make_operand_map(&map_b, bt, dtype, g.m, g.k, BLOCK_N);
  if (rc != MM_OK) return rc;

  const uint32_t tiles = ceil_div(g.n, BLOCK_M) * ceil_div(g.m, BLOCK_N);
  const uint32_t grid = std::min<uint32_t>(tiles, uint32_t(num_sms()));
  const uint32_t k_bytes = uint32_t(size_t(g.k) * eb);
  if (is_f32) {
    auto kern = gemm_tcgen05_kernel<ptx::KIND_TF32, float>;
    MM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
    kern<<<grid, NUM_THREADS, SMEM_BYTES, g.stream>>>(map_a, map_b, static_cast<float *>(g.c), g.n,
                                                     g.m, k_bytes);
  } else {
    auto kern = gemm_tcgen05_kernel<ptx::KIND_F16, __half>;
    MM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
    kern<<<grid, NUM_THREADS, SMEM_BYTES, g.stream>>>(map_a, map_b, static_cast<__half *>(g.c), g.n,
                                                     g.m, k_bytes);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

}  // namespace mm
