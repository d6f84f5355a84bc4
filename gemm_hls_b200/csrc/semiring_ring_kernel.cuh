// Ring variant of the CUDA-core semiring tile kernel for 4-byte element types with A stored row-major:
// BOTH tiles arrive by TMA into a 4-stage ring, stages are handed over through mbarriers (full: TMA
// transaction bytes; empty: one arrival per warp), and there is no block-wide barrier in the main loop.
//
// Why: the inner loop of semiring_tile_kernel alone runs at 44.8 TOp/s for float (Add, Min), the kernel at
// 39.9 (profiles/r01_exp_semiring_issue.jsonl): the difference is the per-k-tile __syncthreads (every 16
// k-steps), the LDG + transposing STS of the A tile and its address arithmetic, and the 8 staging
// registers.  Same arithmetic, same per-element order of operations as semiring_tile_kernel (bit-exact).
//
// Layout.  B tile: [16 k][128 columns], dense, as before.  A tile: [128 rows][16 k] = dense 64-byte rows, as
// it lies in HBM (no transposition); one LDS.128 yields four consecutive k of one row, so the eight rows of a
// thread cost eight LDS.128 per four k-steps — the same count as the k-major tile of semiring_tile_kernel.
// The sixteen lanes of a half-warp share ty, i.e. every 8-lane phase of an LDS.128 reads ONE address
// (broadcast): the A loads are conflict-free without any swizzle; the B loads are as before.
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "ptx_sm100.cuh"
#include "semiring.cuh"
#include "tma_host.cuh"

namespace mm {

struct SemiringRing {
  static constexpr int BM = 128, BN = 128, BK = 16;  // BK elements of 4 bytes = one 64-byte memory word
  static constexpr int STAGES = 4;
  // Tiles are requested AHEAD = STAGES - 2 iterations early: the stage being refilled in iteration kt held
  // tile kt - 2, which every warp released at least one iteration ago — the issuing thread does not have to
  // wait for the slowest warp of the previous tile, so warps may drift by a whole tile.
  static constexpr int AHEAD = STAGES - 2;
  static constexpr int THREADS = 256;
  static constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BK * BN * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 2 * STAGES * 8 + 1024;
};

template <typename T, class Map, class Reduce>
__global__ void __launch_bounds__(256, 2)
semiring_ring_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     T *__restrict__ C, unsigned size_n, unsigned size_k, unsigned size_m) {
  static_assert(sizeof(T) == 4, "ring variant: 4-byte element types");
  using Cfg = SemiringRing;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem0 + STAGES * Cfg::STAGE_BYTES, empty0 = full0 + 8 * STAGES;

  const int tid = threadIdx.x, lane = tid % 32;
  const int tx = tid % 16;  // column quad index
  const int ty = tid / 16;  // row quad index
  const unsigned row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const unsigned k_tiles = size_k / BK;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, Cfg::THREADS / 32);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  // tile kt -> stage kt % STAGES (thread 0 only).  A refill waits until all eight warps have released
  // the stage; see AHEAD for why that wait is normally already satisfied.
  auto load_tile = [&](unsigned kt) {
    const int stage = kt % STAGES;
    if (kt >= STAGES) ptx::mbar_wait(empty0 + 8 * stage, ((kt / STAGES) - 1) & 1);
    const uint32_t as = smem0 + stage * Cfg::STAGE_BYTES, bs = as + Cfg::A_BYTES, bar = full0 + 8 * stage;
    ptx::mbar_arrive_expect_tx(bar, Cfg::STAGE_BYTES);
    ptx::tma_load_2d(as, &tmap_a, bar, int32_t(kt * BK), int32_t(row0), ptx::L2_EVICT_NORMAL);
    ptx::tma_load_2d(bs, &tmap_b, bar, int32_t(col0), int32_t(kt * BK), ptx::L2_EVICT_NORMAL);
  };
  if (tid == 0) {
    for (unsigned kt = 0; kt < unsigned(Cfg::AHEAD) && kt < k_tiles; ++kt) load_tile(kt);
  }

  T acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = Reduce::identity();
  }

  const int r_lo = ty * 4, r_hi = 64 + ty * 4;  // this thread's rows: r_lo .. r_lo + 3 and r_hi .. r_hi + 3

  for (unsigned kt = 0; kt < k_tiles; ++kt) {
    const int stage = kt % STAGES;
    if (tid == 0 && kt + Cfg::AHEAD < k_tiles) load_tile(kt + Cfg::AHEAD);
    ptx::mbar_wait(full0 + 8 * stage, (kt / STAGES) & 1);
    const unsigned char *as = smem_raw + (smem0 - ptx::smem_u32(smem_raw)) + stage * Cfg::STAGE_BYTES;
    const T *bs = reinterpret_cast<const T *>(as + Cfg::A_BYTES);

#pragma unroll
    for (int c = 0; c < BK / 4; ++c) {  // four k per 16-byte chunk of an A row
      T a4[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (i < 4 ? r_lo : r_hi) + (i % 4);
        const Quad<T> q = *reinterpret_cast<const Quad<T> *>(as + r * 64 + c * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) a4[i][v] = q.v[v];
      }
#pragma unroll
      for (int kp = 0; kp < 4; kp += 2) {  // two consecutive k per step, reduced in order (see semiring_tile_kernel)
        const int kk = c * 4 + kp;
        T bf[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const Quad<T> b0 = *reinterpret_cast<const Quad<T> *>(bs + (kk + u) * BN + tx * 4);
          const Quad<T> b1 = *reinterpret_cast<const Quad<T> *>(bs + (kk + u) * BN + 64 + tx * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            bf[u][q] = b0.v[q];
            bf[u][4 + q] = b1.v[q];
          }
        }
        if constexpr (std::is_same<T, float>::value && PackedOp<Map>::value) {
          F32x2 bp[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int p = 0; p < 4; ++p) bp[u][p] = pack_f32x2(bf[u][2 * p], bf[u][2 * p + 1]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const F32x2 a0 = pack_f32x2(a4[i][kp], a4[i][kp]), a1 = pack_f32x2(a4[i][kp + 1], a4[i][kp + 1]);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const F32x2 t0 = PackedOp<Map>::Apply2(a0, bp[0][p]), t1 = PackedOp<Map>::Apply2(a1, bp[1][p]);
              constexpr bool contractable =
                  std::is_same<Map, Product<float>>::value && std::is_same<Reduce, Sum<float>>::value;
              if constexpr (PackedOp<Reduce>::value && !contractable) {
                const F32x2 r = PackedOp<Reduce>::Apply2(
                    PackedOp<Reduce>::Apply2(pack_f32x2(acc[i][2 * p], acc[i][2 * p + 1]), t0), t1);
                unpack_f32x2(r, acc[i][2 * p], acc[i][2 * p + 1]);
              } else {
                float t0l, t0h, t1l, t1h;
                unpack_f32x2(t0, t0l, t0h);
                unpack_f32x2(t1, t1l, t1h);
                acc[i][2 * p] = Reduce::Apply(Reduce::Apply(acc[i][2 * p], t0l), t1l);
                acc[i][2 * p + 1] = Reduce::Apply(Reduce::Apply(acc[i][2 * p + 1], t0h), t1h);
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[i][j] = Reduce::Apply(Reduce::Apply(acc[i][j], Map::Apply(a4[i][kp], bf[0][j])),
                                        Map::Apply(a4[i][kp + 1], bf[1][j]));
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(empty0 + 8 * stage);  // this warp is done reading the stage
  }

  // Write the C tile once, masked to n < N, m < M (the role of WriteC, kernel/Memory.cpp:361-392).
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t row = size_t(row0) + (i / 4) * 64 + ty * 4 + (i % 4);
    if (row >= size_n) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t col = size_t(col0) + h * 64 + tx * 4;
      if (col + 4 <= size_m) {
        Quad<T> out;
#pragma unroll
        for (int q = 0; q < 4; ++q) out.v[q] = acc[i][h * 4 + q];
        *reinterpret_cast<Quad<T> *>(C + row * size_m + col) = out;
      }
    }
  }
}

// Host side: nullptr A = dry run (load the kernel only).  Returns a cudaError_t value as int.
template <typename T, class Map, class Reduce>
int launch_semiring_ring(const void *a, const void *b, void *c, unsigned n, unsigned k, unsigned m,
                         cudaStream_t stream) {
  using Cfg = SemiringRing;
  auto kernel = semiring_ring_kernel<T, Map, Reduce>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES));
  if (e != cudaSuccess || a == nullptr) return static_cast<int>(e);
  CUtensorMap tmap_a, tmap_b;
  // A row-major N x K: box = 128 rows x 16 k (64-byte rows); B row-major K x M: box = 16 k x 128 columns
  if (encode_plain_2d(&tmap_a, a, sizeof(T), n, k, Cfg::BM, Cfg::BK) != 0 ||
      encode_plain_2d(&tmap_b, b, sizeof(T), k, m, Cfg::BK, Cfg::BN) != 0) {
    return static_cast<int>(cudaErrorInvalidValue);
  }
  dim3 grid((m + Cfg::BN - 1) / Cfg::BN, (n + Cfg::BM - 1) / Cfg::BM);
  kernel<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmap_a, tmap_b, static_cast<T *>(c), n, k, m);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace mm
