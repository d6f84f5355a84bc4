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

#include <cmath>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_host.cuh"

namespace mm {
namespace {

using ptx::fence_mbar_init;
using ptx::mbar_arrive;
using ptx::mbar_init;
using ptx::mbar_wait;
using ptx::smem_u32;

constexpr int BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int LDA_S = BK + 4;   // As[BM][LDA_S]   (A row-major tile)
constexpr int LDB_S = BN + 4;   // Bs[BK][LDB_S]
constexpr int B_TILE = BK * LDB_S;
// BM (rows of C per CTA) is 128, or 64 when that fills the last wave better (see launch_dmma)
template <int BM>
struct Tile {
  static constexpr int LDAT_S = BM + 4;  // AsT[BK][LDAT_S] (A stored K x N)
  static constexpr int A_TILE = (BM * LDA_S > BK * LDAT_S) ? BM * LDA_S : BK * LDAT_S;
  static constexpr size_t SMEM_BYTES = size_t(STAGES) * (A_TILE + B_TILE) * sizeof(double);
};

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
template <bool TRANSPOSED_A, int BM, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32, 1)
gemm_dmma_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
                 unsigned size_n, unsigned size_k, unsigned size_m) {
  constexpr int LDAT_S = Tile<BM>::LDAT_S;
  constexpr int A_TILE = Tile<BM>::A_TILE;
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
      // BK k-rows x BM n-cols; BM / 2 chunks per row
#pragma unroll
      for (int i = 0; i < BK * BM / 2 / THREADS; ++i) {
        const int c = tid + i * THREADS;
        const int kk = c / (BM / 2), part = c % (BM / 2);
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


// ---- warp-specialised variant --------------------------------------------------------------------
// In the kernel above every warp issues its share of the cp.async prefetch between the block
// barrier and its first DMMA of a k-tile; ncu attributes ~12 % of the warp time to that sequence
// (LDGSTS operand-read scoreboards, barrier skew, first LDS) and the DMMA pipe idles meanwhile
// (88 % active).  Here one extra producer warp owns all global -> shared traffic and the WM x WN
// compute warps never leave the LDS / DMMA loop: stages are handed over through mbarriers
// (full[s]: completion of the producer's cp.asyncs via cp.async.mbarrier.arrive.noinc;
//  empty[s]: one arrival per compute warp), so compute warps drift against each other instead of
// meeting at a block-wide barrier every k-tile.
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <bool TRANSPOSED_A, int BM, int WM, int WN, int PW>
__global__ void __launch_bounds__((WM * WN + PW) * 32, 1)
gemm_dmma_ws_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
                    unsigned size_n, unsigned size_k, unsigned size_m, int dbg) {
  constexpr int LDAT_S = Tile<BM>::LDAT_S;
  constexpr int A_TILE = Tile<BM>::A_TILE;
  constexpr int NCW = WM * WN;  // compute warps; warps NCW .. NCW + PW - 1 are the producers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *As = reinterpret_cast<double *>(smem_raw);
  double *Bs = As + STAGES * A_TILE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + STAGES * B_TILE);  // full[STAGES], empty[STAGES]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);

  constexpr int MI = BM / (WM * 8);
  constexpr int NJ = BN / (WN * 8);
  const int tid = threadIdx.x;
  const int warp = tid / 32, lane = tid % 32;
  const size_t row0 = size_t(blockIdx.y) * BM;
  const size_t col0 = size_t(blockIdx.x) * BN;
  const unsigned k_tiles = (size_k + BK - 1) / BK;
  const size_t lrow0 = (dbg & 2) ? 0 : row0, lcol0 = (dbg & 2) ? 0 : col0;  // TEMP experiment

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 32 * PW);
      mbar_init(empty0 + 8 * s, NCW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= NCW) {
    // ------------------------------ producer warps ------------------------------
    // A single warp cannot issue a k-tile's 128 LDGSTS in one k-tile time (each holds its address
    // registers until the LSU has consumed them), so the rows are dealt out to PW warps.
    const int pw = warp - NCW;
    for (unsigned kt = 0; kt < k_tiles; ++kt) {
      const int stage = kt % STAGES;
      if (kt >= STAGES) mbar_wait(empty0 + 8 * stage, ((kt / STAGES) - 1) & 1);
      const unsigned k0 = (dbg & 2) ? 0 : kt * BK;
      double *as = As + stage * A_TILE;
      double *bs = Bs + stage * B_TILE;
      if (dbg & 1) {
        cp_async_arrive_noinc(full0 + 8 * stage);
        continue;
      }
      if (!TRANSPOSED_A) {
        // BM rows x 16 chunks of 16 bytes; a warp instruction covers two rows
        const int part = lane % (BK / 2), rsub = lane / (BK / 2);
        const unsigned kk = k0 + part * 2;
        const bool valid = kk < size_k;  // K % 8 == 0: a chunk is entirely in or out
        const double *src = A + (valid ? kk : 0);
#pragma unroll 8
        for (int r = rsub + 2 * pw; r < BM; r += 2 * PW) {
          size_t row = lrow0 + r;
          if (row >= size_n) row = size_n - 1;
          cp_async16(as + r * LDA_S + part * 2, src + row * size_k, valid);
        }
      } else {
        // BK k-rows x BM / 2 chunks
#pragma unroll 4
        for (int kk = pw; kk < BK; kk += PW) {
          const bool valid = (k0 + kk) < size_k;
          const double *src = A + size_t(valid ? k0 + kk : 0) * size_n;
#pragma unroll
          for (int part = lane; part < BM / 2; part += 32) {
            size_t n = lrow0 + part * 2;
            if (n + 2 > size_n) n = size_n - 2;  // N % 2 == 0 checked by the launcher
            cp_async16(as + kk * LDAT_S + part * 2, src + n, valid);
          }
        }
      }
#pragma unroll 4
      for (int kk = pw; kk < BK; kk += PW) {
        const bool valid = (k0 + kk) < size_k;
        const double *src = B + size_t(valid ? k0 + kk : 0) * size_m;
#pragma unroll
        for (int part = lane; part < BN / 2; part += 32) {
          size_t col = lcol0 + part * 2;
          if (col + 2 > size_m) col = size_m - 2;
          cp_async16(bs + kk * LDB_S + part * 2, src + col, valid);
        }
      }
      cp_async_arrive_noinc(full0 + 8 * stage);  // arrives once this lane's copies have landed
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // nothing in flight when the CTA retires
    return;
  }

  // ------------------------------ compute warps ------------------------------
  const int wr = warp / WN, wc = warp % WN;
  const int g = lane / 4, q = lane % 4;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (unsigned kt = 0; kt < k_tiles; ++kt) {
    const int stage = kt % STAGES;
    mbar_wait(full0 + 8 * stage, (kt / STAGES) & 1);
    const double *as = As + stage * A_TILE;
    const double *bs = Bs + stage * B_TILE;
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
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * stage);  // this warp is done reading the stage
  }

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

template <int BM, int WM, int WN, int PW>
static int launch_dmma_ws(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  constexpr size_t SMEM = Tile<BM>::SMEM_BYTES + 2 * STAGES * sizeof(uint64_t);
  constexpr int THREADS = (WM * WN + PW) * 32;
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_ws_kernel<false, BM, WM, WN, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_ws_kernel<true, BM, WM, WN, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  if (g.dry_run) return MM_OK;
  static const int dbg = [] { const char *e = std::getenv("MM_DMMA_DEBUG"); return e ? std::atoi(e) : 0; }();  // TEMP
  dim3 grid(ceil_div(g.m, BN), ceil_div(g.n, BM));
  const double *a = static_cast<const double *>(g.a);
  const double *b = static_cast<const double *>(g.b);
  double *c = static_cast<double *>(g.c);
  if (ta) {
    gemm_dmma_ws_kernel<true, BM, WM, WN, PW><<<grid, THREADS, SMEM, g.stream>>>(a, b, c, g.n, g.k, g.m, dbg);
  } else {
    gemm_dmma_ws_kernel<false, BM, WM, WN, PW><<<grid, THREADS, SMEM, g.stream>>>(a, b, c, g.n, g.k, g.m, dbg);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}


// ---- TMA-fed variant -----------------------------------------------------------------------------
// Even four LDGSTS producer warps cost the DMMA pipe ~4 % (36.5 TF/s with the loads switched off,
// 35.1 with them).  Here one thread feeds the ring with cp.async.bulk.tensor: no LSU instructions,
// no address arithmetic, out-of-range rows / columns / k zero-filled by the TMA unit.
// TMA writes dense tiles, so bank conflicts are avoided by the 128-byte swizzle (16-byte chunk
// index XOR (tile row % 8)) plus a permutation of which physical rows / columns the eight row- or
// column-slots g of an m8n8k4 fragment stand for (k stays natural: step s, slot q <-> k = 4 s + q):
//   * A row-major: tiles of [BM rows][16 k]; accumulator tile i, slot g <-> row 2 g + (i % 2) + 16 (i / 2):
//     the four rows of a half-warp have row % 8 = {0,2,4,6} (+ i % 2), which XORs the two chunks a
//     row's four k-slots touch onto four disjoint aligned chunk pairs
//   * B (and A when stored K x N): tiles of [32 k][16 columns]; tile j, slot g <-> column
//     (g % 2) + 2 (g / 4) + 8 ((g / 2) % 2) + 4 (j % 2) + 16 (j / 2): a half-warp touches chunks c and
//     c ^ 4, which stay disjoint under the XOR with k % 8 = 4 (s % 2) + q
// Every half-warp LDS.64 then reads sixteen distinct 8-byte words of one 128-byte bank row.
constexpr int TMA_WM = 2, TMA_WN = 4;

template <bool TRANSPOSED_A, int BM>
__global__ void __launch_bounds__((TMA_WM * TMA_WN + 1) * 32, 1)
gemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     double *__restrict__ C, unsigned size_n, unsigned size_k, unsigned size_m) {
  constexpr int WM = TMA_WM, WN = TMA_WN, NCW = WM * WN;
  constexpr int MI = BM / (WM * 8), NJ = BN / (WN * 8);
  constexpr int WROWS = BM / WM;  // rows of C per warp
  constexpr uint32_t A_BYTES = BM * BK * 8, B_BYTES = BK * BN * 8, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  const uint32_t full0 = smem0 + STAGES * STAGE_BYTES, empty0 = full0 + 8 * STAGES;

  const int tid = threadIdx.x;
  const int warp = tid / 32, lane = tid % 32;
  const unsigned row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const unsigned k_tiles = (size_k + BK - 1) / BK;

  if (tid == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, NCW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == NCW) {
    if (lane != 0) return;
    for (unsigned kt = 0; kt < k_tiles; ++kt) {
      const int stage = kt % STAGES;
      if (kt >= STAGES) mbar_wait(empty0 + 8 * stage, ((kt / STAGES) - 1) & 1);
      const uint32_t as = smem0 + stage * STAGE_BYTES, bs = as + A_BYTES, bar = full0 + 8 * stage;
      const int k0 = int(kt * BK);
      ptx::mbar_arrive_expect_tx(bar, STAGE_BYTES);
      if (!TRANSPOSED_A) {
#pragma unroll
        for (int kh = 0; kh < BK / 16; ++kh)
          ptx::tma_load_2d(as + kh * BM * 128, &map_a, bar, k0 + kh * 16, int(row0), ptx::L2_EVICT_NORMAL);
      } else {
#pragma unroll
        for (int sl = 0; sl < BM / 16; ++sl)
          ptx::tma_load_2d(as + sl * 4096, &map_a, bar, int(row0) + sl * 16, k0, ptx::L2_EVICT_NORMAL);
      }
#pragma unroll
      for (int sl = 0; sl < BN / 16; ++sl)
        ptx::tma_load_2d(bs + sl * 4096, &map_b, bar, int(col0) + sl * 16, k0, ptx::L2_EVICT_NORMAL);
    }
    return;
  }

  const int wr = warp / WN, wc = warp % WN;
  const int g = lane / 4, q = lane % 4;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // slot permutation within a 16-wide tile (see above); t16(x) for x = i or j
  const int perm16 = (g % 2) + 2 * (g / 4) + 8 * ((g / 2) % 2);
  // physical row (within the CTA tile) of accumulator tile i, row-slot g
  auto c_row = [&](int i) {
    return TRANSPOSED_A ? wr * WROWS + 16 * (i / 2) + 4 * (i % 2) + perm16
                        : wr * WROWS + 16 * (i / 2) + (i % 2) + 2 * g;
  };
  auto lds = [](uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  };

  // Fragment addresses = (stage base + per-thread base + compile-time tile offset) XOR a compile-time
  // chunk constant: everything below bit 7 of the per-thread bases is the thread's own swizzle term,
  // all tile offsets are multiples of 128, so the 16-byte-chunk XOR can be applied to the sum.
  const uint32_t sw_base = q * 128 + (((perm16 / 2) ^ q) * 16) + (perm16 % 2) * 8;  // [32 k][16] tiles
  const uint32_t a_base = TRANSPOSED_A ? (wr * WROWS / 16) * 4096 + sw_base
                                       : (wr * WROWS + 2 * g) * 128 + (((q / 2) ^ ((2 * g) % 8)) * 16) + (q % 2) * 8;
  const uint32_t b_base = wc * (NJ / 2) * 4096 + sw_base;

  for (unsigned kt = 0; kt < k_tiles; ++kt) {
    const int stage = kt % STAGES;
    mbar_wait(full0 + 8 * stage, (kt / STAGES) & 1);
    const uint32_t as = smem0 + stage * STAGE_BYTES + a_base, bs = smem0 + stage * STAGE_BYTES + A_BYTES + b_base;
#pragma unroll
    for (int s = 0; s < BK / 4; ++s) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        if (!TRANSPOSED_A) {
          af[i] = lds((as + (s / 4) * (BM * 128) + (16 * (i / 2) + (i % 2)) * 128) ^ (((2 * (s % 4)) ^ (i % 2)) * 16));
        } else {
          af[i] = lds((as + (i / 2) * 4096 + s * 512) ^ (((2 * (i % 2)) ^ (4 * (s % 2))) * 16));
        }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = lds((bs + (j / 2) * 4096 + s * 512) ^ (((2 * (j % 2)) ^ (4 * (s % 2))) * 16));
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * stage);
  }

#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const size_t row = size_t(row0) + c_row(i);
    if (row >= size_n) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      // column-slots 2q, 2q+1 of tile j (same permutation as the B fragment)
      const size_t col = size_t(col0) + wc * (NJ * 8) + 16 * (j / 2) + 4 * (j % 2) + 2 * (q / 2) + 8 * (q % 2);
      if (col + 2 <= size_m) {
        *reinterpret_cast<double2 *>(C + row * size_m + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

template <int BM>
static int launch_dmma_tma(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  constexpr size_t SMEM = size_t(STAGES) * (BM * BK + BK * BN) * 8 + 2 * STAGES * 8 + 1024;
  constexpr int THREADS = (TMA_WM * TMA_WN + 1) * 32;
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_tma_kernel<false, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_tma_kernel<true, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  if (g.dry_run) return MM_OK;
  CUtensorMap map_a, map_b;
  const int ra = ta ? encode_sw128_2d_f64(&map_a, g.a, g.k, g.n, BK) : encode_sw128_2d_f64(&map_a, g.a, g.n, g.k, BM);
  const int rb = encode_sw128_2d_f64(&map_b, g.b, g.k, g.m, BK);
  if (ra != 0 || rb != 0) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled failed for the f64 operands");
  dim3 grid(ceil_div(g.m, BN), ceil_div(g.n, BM));
  double *c = static_cast<double *>(g.c);
  if (ta) {
    gemm_dmma_tma_kernel<true, BM><<<grid, THREADS, SMEM, g.stream>>>(map_a, map_b, c, g.n, g.k, g.m);
  } else {
    gemm_dmma_tma_kernel<false, BM><<<grid, THREADS, SMEM, g.stream>>>(map_a, map_b, c, g.n, g.k, g.m);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

template <int BM, int WM, int WN>
static int launch_dmma_variant(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  constexpr size_t SMEM = Tile<BM>::SMEM_BYTES;
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<false, BM, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  MM_CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<true, BM, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
  if (g.dry_run) return MM_OK;
  dim3 grid(ceil_div(g.m, BN), ceil_div(g.n, BM));
  const double *a = static_cast<const double *>(g.a);
  const double *b = static_cast<const double *>(g.b);
  double *c = static_cast<double *>(g.c);
  if (ta) {
    gemm_dmma_kernel<true, BM, WM, WN><<<grid, WM * WN * 32, SMEM, g.stream>>>(a, b, c, g.n, g.k, g.m);
  } else {
    gemm_dmma_kernel<false, BM, WM, WN><<<grid, WM * WN * 32, SMEM, g.stream>>>(a, b, c, g.n, g.k, g.m);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

}  // namespace

int launch_dmma(const GemmArgs &g) {
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  if (ta && (g.n % 2 != 0)) return fail(MM_ERR_UNSUPPORTED, "dmma path with transposed A needs even N");
  // Tile height: 128 rows per CTA, or 64 when the 128-row tiling leaves the last wave mostly empty
  // (e.g. a 1024-row block of the 8-GPU split of 8192^3: 512 tiles on 148 SMs = 3.46 waves; 1024
  // half-height tiles = 6.92 waves of half the duration).  The half-height tile re-reads B twice
  // as often per output row, so it has to win by more than 5 % to be chosen.
  // MM_DMMA_TILE_ROWS=64|128 forces one of them; MM_DMMA_WARPS=8 selects the 2 x 4 warp layout.
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const double t128 = double(ceil_div(g.n, 128)) * ceil_div(g.m, BN), t64 = double(ceil_div(g.n, 64)) * ceil_div(g.m, BN);
  const double cost128 = std::ceil(t128 / sms), cost64 = 0.5 * 1.05 * std::ceil(t64 / sms);
  static const int forced = [] {
    const char *e = std::getenv("MM_DMMA_TILE_ROWS");
    return e ? std::atoi(e) : 0;
  }();
  static const int warps = [] {
    const char *e = std::getenv("MM_DMMA_WARPS");
    return (e && std::atoi(e) == 8) ? 8 : 16;
  }();
  const bool use64 = forced == 64 || (forced != 128 && cost64 < cost128);
  static const int ws = [] {
    const char *e = std::getenv("MM_DMMA_WS");
    return e ? std::atoi(e) : 1;
  }();
  static const int use_tma = [] {
    const char *e = std::getenv("MM_DMMA_TMA");
    return e ? std::atoi(e) : 1;
  }();
  const bool aligned = (reinterpret_cast<uintptr_t>(g.a) % 16 == 0) && (reinterpret_cast<uintptr_t>(g.b) % 16 == 0);
  if (use_tma && aligned && get_encode_fn()) return use64 ? launch_dmma_tma<64>(g) : launch_dmma_tma<128>(g);
  static const int pw = [] {
    const char *e = std::getenv("MM_DMMA_PRODUCERS");
    return e ? std::atoi(e) : 4;
  }();
  if (ws && pw == 1) return use64 ? launch_dmma_ws<64, 2, 4, 1>(g) : launch_dmma_ws<128, 2, 4, 1>(g);
  if (ws && pw == 2) return use64 ? launch_dmma_ws<64, 2, 4, 2>(g) : launch_dmma_ws<128, 2, 4, 2>(g);
  if (ws) return use64 ? launch_dmma_ws<64, 2, 4, 4>(g) : launch_dmma_ws<128, 2, 4, 4>(g);
  if (use64) return launch_dmma_variant<64, 2, 8>(g);
  return warps == 8 ? launch_dmma_variant<128, 2, 4>(g) : launch_dmma_variant<128, 4, 4>(g);
}

}  // namespace mm
