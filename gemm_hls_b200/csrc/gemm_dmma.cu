// Tensor-core path for the dense (Multiply, Add) contraction on double:  C = A * B  in FP64.
// tcgen05 has no f64 kind, so this is the warp-level DMMA instruction
// mma.sync.aligned.m8n8k4.row.col.f64 fed from a cp.async shared-memory ring.
// B200 counterpart of the reference's PE chain for MM_DATA_TYPE=double
// (kernel/Compute.cpp:53-146; README.md:8 quotes 132 GFLOP/s for it on a VCU1525).
//
// CTA tile 128 x 128, BK = 32, 256 threads = 8 warps as 2 (rows) x 4 (cols), warp tile 64 x 32 =
// 8 x 4 m8n8 accumulator tiles (64 doubles per thread).  A and B tiles keep their global
// orientation in shared memory; row pitches are padded by 4 doubles so that the 8-byte fragment
// reads of each half-warp hit 16 distinct 8-byte banks.
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace mm {
namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int LDA_S = BK + 4;   // As[BM][LDA_S]   (A row-major tile)
constexpr int LDAT_S = BM + 4;  // AsT[BK][LDAT_S] (A stored K x N)
constexpr int LDB_S = BN + 4;   // Bs[BK][LDB_S]
constexpr int A_TILE = (BM * LDA_S > BK * LDAT_S) ? BM * LDA_S : BK * LDAT_S;
constexpr int B_TILE = BK * LDB_S;
constexpr size_t SMEM_BYTES = size_t(STAGES) * (A_TILE + B_TILE) * sizeof(double);

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int src_bytes = valid ? 16 : 0;  // 0 -> zero-fill (neutral for (Multiply, Add))
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// WM x WN warps; each warp owns a (BM / WM) x (BN / WN) block of C as MI x NJ m8n8 accumulator tiles.
// 2 x 4 warps (64 x 32 per warp, 64 accumulators per thread) minimises fragment loads per DMMA;
// 4 x 4 warps (32 x 32 per warp, 32 accumulators) doubles the warps per scheduler.
template <bool TRANSPOSED_A, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32, 1)
gemm_dmma_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
                 unsigned size_n, unsigned size_k, unsigned size_m) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *As = reinterpret_cast<double *>(smem_raw);
  double *Bs = As + STAGES * A_TILE;

  constexpr int THREADS = WM * WN * 32;
  constexpr int MI = BM / (WM * 8);  // m8n8 tiles per warp along M
  constexpr int NJ = BN / (WN * 8);  // ... along N
  const int tid = threadIdx.x;
  const int warp = tid / 32, lane = tid % 32;
  const int wr = warp / WN;  // warp row    -> rows [wr * MI * 8, +MI * 8)
  const int wc = warp % WN;  // warp column -> cols [wc * NJ * 8, +NJ * 8)
  const int g = lane / 4;   // fragment row / col within an 8-wide tile
  const int q = lane % 4;   // fragment k index
  const size_t row0 = size_t(blockIdx.y) * BM;
  const size_t col0 = size_t(blockIdx.x) * BN;

  auto load_tile = [&](int stage, unsigned k0) {
    double *as = As + stage * A_TILE;
    double *bs = Bs + stage * B_TILE;
    if (!TRANSPOSED_A) {
      // 128 rows x BK doubles as 16-byte chunks; BK / 2 chunks per row
#pragma unroll
      for (int i = 0; i < BM * BK / 2 / THREADS; ++i) {
        const int c = tid + i * THREADS;
        const int r = c / (BK / 2), part = c % (BK / 2);
        size_t row = row0 + r;
        if (row >= size_n) row = size_n - 1;
        const unsigned kk = k0 + part * 2;
        const bool valid = kk < size_k;  // K % 8 == 0: a chunk is entirely in or out
        cp_async16(as + r * LDA_S + part * 2, A + row * size_k + (valid ? kk : 0), valid);
      }
    } else {
      // BK k-rows x 128 n-cols; 64 chunks per row
#pragma unroll
      for (int i = 0; i < BK * BM / 2 / THREADS; ++i) {
        const int c = tid + i * THREADS;
        const int kk = c / 64, part = c % 64;
        size_t n = row0 + part * 2;
        if (n + 2 > size_n) n = size_n - 2;  // N % 2 == 0 checked by the launcher
        const bool valid = (k0 + kk) < size_k;
        cp_async16(as + kk * LDAT_S + part * 2, A + size_t(valid ? k0 + kk : 0) * size_n + n, valid);
      }
    }
#pragma unroll
    for (int i = 0; i < BK * BN / 2 / THREADS; ++i) {
      const int c = tid + i * THREADS;
      const int kk = c / 64, part = c % 64;
      size_t col = col0 + part * 2;
      if (col + 2 > size_m) col = size_m - 2;
      const bool valid = (k0 + kk) < size_k;
      cp_async16(bs + kk * LDB_S + part * 2, B + size_t(valid ? k0 + kk : 0) * size_m + col, valid);
    }
  };

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const unsigned k_tiles = (size_k + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (unsigned(s) < k_tiles) load_tile(s, s * BK);
    cp_async_commit();
  }

  for (unsigned kt = 0; kt < k_tiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // prefetch tile kt + STAGES - 1 into the stage freed by iteration kt - 1
    const unsigned nk = kt + STAGES - 1;
    if (nk < k_tiles) load_tile(nk % STAGES, nk * BK);
    cp_async_commit();

    const double *as = As + (kt % STAGES) * A_TILE;
    const double *bs = Bs + (kt % STAGES) * B_TILE;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = wr * MI * 8 + i * 8 + g;
        af[i] = TRANSPOSED_A ? as[(k4 + q) * LDAT_S + r] : as[r * LDA_S + k4 + q];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = bs[(k4 + q) * LDB_S + wc * NJ * 8 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // C fragment of m8n8: thread holds (row g, cols 2q, 2q+1)
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const size_t row = row0 + wr * MI * 8 + i * 8 + g;
    if (row >= size_n) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const size_t col = col0 + wc * NJ * 8 + j * 8 + q * 2;
      if (col + 2 <= size_m) {
        *reinterpret_cast<double2 *>(C + row * size_m + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

template <int WM, int WN>
static int launch_dmma_variant(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<false, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<true, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
  if (g.dry_run) return MM_OK;
  dim3 grid(ceil_div(g.m, BN), ceil_div(g.n, BM));
  const double *a = static_cast<const double *>(g.a);
  const double *b = static_cast<const double *>(g.b);
  double *c = static_cast<double *>(g.c);
  if (ta) {
    gemm_dmma_kernel<true, WM, WN><<<grid, WM * WN * 32, SMEM_BYTES, g.stream>>>(a, b, c, g.n, g.k, g.m);
  } else {
    gemm_dmma_kernel<false, WM, WN><<<grid, WM * WN * 32, SMEM_BYTES, g.stream>>>(a, b, c, g.n, g.k, g.m);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

}  // namespace

int launch_dmma(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  if (ta && (g.n % 2 != 0)) return fail(MM_ERR_UNSUPPORTED, "dmma path with transposed A needs even N");
  // MM_DMMA_WARPS=8 selects the 2 x 4 warp layout for A/B measurements (default 16 = 4 x 4)
  static const int warps = [] {
    const char *e = std::getenv("MM_DMMA_WARPS");
    return (e && std::atoi(e) == 8) ? 8 : 16;
  }();
  return warps == 8 ? launch_dmma_variant<2, 4>(g) : launch_dmma_variant<4, 4>(g);
}

}  // namespace mm
