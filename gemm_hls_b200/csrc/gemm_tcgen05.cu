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
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_host.cuh"

namespace mm {
namespace {

constexpr int BLOCK_M = 128;          // C rows per CTA      (UMMA M = 128 * CTA group size)
constexpr int BLOCK_N = 256;          // C cols per tile     (UMMA N)
constexpr int BLOCK_K_BYTES = 128;    // one 128-byte swizzle atom of K per stage
constexpr int UMMA_K_BYTES = 32;      // K extent of one tcgen05.mma
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512
constexpr int NUM_THREADS = 192;
constexpr int MN_ATOM = 64;                       // elements of N per MN-major swizzle atom (128 B of f16)
constexpr int MN_ATOM_BYTES = 64 * BLOCK_K_BYTES; // one atom: BLOCK_K (= 64 for f16) k-rows x 128 B
constexpr int RASTER_GROUP_ROWS = 2048;  // C rows per rasterisation group (L2 reuse of B^T panels)

// Per-variant geometry.  CG = 1: one CTA computes a 128 x 256 tile and stages A (128 rows) + B^T
// (256 rows) per k-block.  CG = 2 (cta_group::2): a CTA PAIR computes 256 x 256 with ONE
// tcgen05.mma per k-step issued by the leader CTA; each CTA stages only its own 128 A rows and its
// half (128 rows) of the B^T tile, so per-SM shared-memory fill and L2 traffic drop by a third and
// the freed shared memory deepens the ring from 4 to 6 stages.
template <int CG>
struct Geo {
  static constexpr int LOAD_N = BLOCK_N / CG;                       // B^T rows staged per CTA
  static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K_BYTES;     // 16 KiB
  static constexpr int B_STAGE_BYTES = LOAD_N * BLOCK_K_BYTES;      // 32 / 16 KiB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;
  static constexpr int TILE_ROWS = BLOCK_M * CG;                    // C rows per CTA group
  static constexpr int RASTER_GROUP = RASTER_GROUP_ROWS / TILE_ROWS;
  static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct TileCoord {
  uint32_t r, c;
};

// Grouped rasterisation: RASTER_GROUP row-tiles sweep all column-tiles together so that the
// concurrently running tiles share A row-panels and B column-panels through L2.
__device__ __forceinline__ TileCoord tile_coord(uint32_t t, uint32_t tiles_r, uint32_t tiles_c,
                                                uint32_t raster_group) {
  const uint32_t per_group = raster_group * tiles_c;
  const uint32_t g = t / per_group;
  const uint32_t first = g * raster_group;
  const uint32_t gsize = min(raster_group, tiles_r - first);
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

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// C[rows x cols] = A'[rows x k] * Bt[cols x k]^T ; A', Bt K-major, described by the tensor maps.
// CG == 2 must be launched with cluster dimension (2, 1, 1).
// BMN: the B operand is read MN-major straight from the reference's row-major B (K x M) — no
// transposed copy.  Implemented for kind::f16 (64-element atoms); kind::tf32 needs B rounded to
// TF32 anyway, so its prepared copy is written K-major.
// FUSE_A (float only): the TF32 rounding of A runs INSIDE this kernel on four extra warps of every
// CTA.  They sweep A in row order in granules of PREP_ROWS rows (all CTAs share each granule), write
// the rounded copy that tmap_a describes, and count finished granules in `fuse.a_done`; the TMA
// producer starts a tile only when its granule is complete.  Only the first granules are exposed
// (the GEMM consumes 2048 rows per ~1.2 ms, the sweep rounds the whole of A in ~0.5 ms).
struct FuseA {
  const float4 *a_raw;    // caller's A (row-major rows x k)
  float4 *a_prep;         // rounded copy (what tmap_a points at)
  unsigned int *a_done;   // one counter per granule, zeroed by the launcher; complete == gridDim.x
  uint32_t k_elems;
  uint32_t late_warps;    // rounding warps per CTA that keep working after the first raster group (1..4)
  uint32_t pre_done;      // leading granules already rounded by a separate launch before this kernel
};
constexpr int PREP_WARPS = 4;
constexpr int PREP_ROWS = 256;   // granule height = pair-tile height, so a CTA's 128 rows lie in one granule

template <int KIND, typename TOut, int CG, bool BMN, bool FUSE_A>
__global__ void __launch_bounds__(NUM_THREADS + (FUSE_A ? PREP_WARPS * 32 : 0), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, TOut *__restrict__ C, uint32_t rows,
                    uint32_t cols, uint32_t k_bytes, uint32_t num_stages, uint32_t raster_group,
                    uint64_t l2_policy, unsigned int *tile_sync, FuseA fuse, unsigned long long *dbg) {
  using G = Geo<CG>;
  const int STAGES = int(num_stages);  // <= G::STAGES (what the shared-memory allocation holds)
  extern __shared__ unsigned char smem_raw[];
  // 128B-swizzled tiles must start on a 1024-byte boundary (same offset in both CTAs of a pair).
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + STAGES * G::A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + STAGES * G::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  const uint32_t group_id = blockIdx.x / CG;
  const uint32_t num_groups = gridDim.x / CG;

  const uint32_t tiles_r = (rows + G::TILE_ROWS - 1) / G::TILE_ROWS;
  const uint32_t tiles_c = (cols + BLOCK_N - 1) / BLOCK_N;
  const uint32_t num_tiles = tiles_r * tiles_c;
  const uint32_t num_kb = (k_bytes + BLOCK_K_BYTES - 1) / BLOCK_K_BYTES;
  constexpr int ELEM_BYTES = (KIND == ptx::KIND_TF32) ? 4 : 2;
  constexpr int BLOCK_K_ELEMS = BLOCK_K_BYTES / ELEM_BYTES;

  if (CG == 2) cluster_sync_all();  // both CTAs resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      // ONE arrival: the (leader's) producer's arrive.expect_tx, which books the bytes of BOTH CTAs.
      // The peer's TMA transactions complete_tx on the leader's barrier; no peer arrival is needed
      // (a remote mbarrier.arrive per k-block stalled the peer's producer for ~400 cycles: measured
      // with MM_TCGEN05_DEBUG, the MMA thread waited 68 % of the time on data).  The phase cannot
      // complete early because the expected byte count includes the peer's half, and the peer
      // cannot run a phase ahead because its empty barrier is released by the same tcgen05.commit.
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);  // tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);        // tcgen05.commit
      ptx::mbar_init(tmem_empty_bar(s), 4 * CG);  // one arrive per epilogue warp of the group (leader's copy)
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<CG>(tmem_slot, TMEM_COLS);
  }
  ptx::tcgen05_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  ptx::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer (one per CTA) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long wait_empty = 0, t_begin = clock64();
      uint32_t tile_iter = 0;
      int32_t a_ready = -1;
      for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++tile_iter) {
        // Soft wave barrier: do not start fetching tile #j before every CTA group has finished
        // fetching its tile #(j-1).  A ring deep enough to hide DRAM latency removes the L2-miss
        // back-pressure that otherwise keeps the co-running tiles in lock-step, the groups drift,
        // and the A / B^T panels they share are re-read from DRAM (measured: 17.7 -> 32 GB at
        // 16384^3).  Purely a performance hint: the wait is bounded, correctness never depends on it.
        if (tile_sync != nullptr && tile_iter > 0) {
          const uint32_t target = min(tile_iter * num_groups, num_tiles);
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile unsigned int *>(tile_sync) < target) {
            if (clock64() - t0 > 100000) break;  // ~50 us: give up, stay correct
          }
        }
        const TileCoord tc = tile_coord(t, tiles_r, tiles_c, raster_group);
        const int32_t a_row = tc.r * G::TILE_ROWS + cta_rank * BLOCK_M;
        const int32_t b_row = tc.c * BLOCK_N + cta_rank * G::LOAD_N;
        if (FUSE_A) {
          // rows [a_row, a_row + 128) lie in granule a_row / PREP_ROWS; granules complete in order
          const int32_t granule = a_row / PREP_ROWS;
          if (granule > a_ready && granule >= int32_t(fuse.pre_done)) {
            // All CTAs of this persistent grid are resident (grid <= SMs, 1 CTA per SM), so the
            // rounding warps of every CTA make progress while this thread spins.  The bound only
            // turns an impossible-to-satisfy wait (e.g. a tool that serialises CTAs) into a trap.
            const volatile unsigned int *flag = fuse.a_done + granule;
            const long long t0 = clock64();
            while (*flag < gridDim.x) {
              if (clock64() - t0 > (1ll << 33)) __trap();
            }
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy writes -> TMA reads
            a_ready = granule;
          }
        }
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          if (dbg) {
            const long long t0 = clock64();
            ptx::mbar_wait(empty_bar(stage), phase ^ 1);
            wait_empty += clock64() - t0;
          } else {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          }
          if (CG == 1) {
            ptx::mbar_arrive_expect_tx(full_bar(stage), G::STAGE_BYTES);
            ptx::tma_load_2d(smem_a0 + stage * G::A_STAGE_BYTES, &tmap_a, full_bar(stage),
                             kb * BLOCK_K_ELEMS, a_row, l2_policy);
            if (BMN) {
#pragma unroll
              for (int j = 0; j < G::LOAD_N / MN_ATOM; ++j) {
                ptx::tma_load_2d(smem_b0 + stage * G::B_STAGE_BYTES + j * MN_ATOM_BYTES, &tmap_b, full_bar(stage),
                                 b_row + j * MN_ATOM, kb * BLOCK_K_ELEMS, l2_policy);
              }
            } else {
              ptx::tma_load_2d(smem_b0 + stage * G::B_STAGE_BYTES, &tmap_b, full_bar(stage),
                               kb * BLOCK_K_ELEMS, b_row, l2_policy);
            }
          } else {
            // both CTAs' bytes are accounted on the LEADER's barrier (peer bit 24 cleared)
            const uint32_t leader_bar = full_bar(stage) & 0xFEFFFFFFu;
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * G::STAGE_BYTES);
            ptx::tma_load_2d_2sm(smem_a0 + stage * G::A_STAGE_BYTES, &tmap_a, leader_bar,
                                 kb * BLOCK_K_ELEMS, a_row, l2_policy);
            if (BMN) {
#pragma unroll
              for (int j = 0; j < G::LOAD_N / MN_ATOM; ++j) {
                ptx::tma_load_2d_2sm(smem_b0 + stage * G::B_STAGE_BYTES + j * MN_ATOM_BYTES, &tmap_b, leader_bar,
                                     b_row + j * MN_ATOM, kb * BLOCK_K_ELEMS, l2_policy);
              }
            } else {
              ptx::tma_load_2d_2sm(smem_b0 + stage * G::B_STAGE_BYTES, &tmap_b, leader_bar,
                                   kb * BLOCK_K_ELEMS, b_row, l2_policy);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tile_sync != nullptr && cta_rank == 0) atomicAdd(tile_sync, 1u);  // this group fetched its tile
      }
      if (dbg) {
        dbg[blockIdx.x * 8 + 0] = wait_empty;
        dbg[blockIdx.x * 8 + 1] = clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(KIND, BLOCK_M * CG, BLOCK_N, BMN);
      uint32_t stage = 0, phase = 0, iter = 0;
      long long wait_full = 0, wait_tmem = 0, t_begin = clock64();
      for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++iter) {
        const uint32_t as = iter & 1u;
        const uint32_t aphase = (iter >> 1) & 1u;
        if (dbg) {
          const long long t0 = clock64();
          ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1);
          wait_tmem += clock64() - t0;
        } else {
          ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        }
        ptx::tcgen05_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          if (dbg) {
            const long long t0 = clock64();
            ptx::mbar_wait(full_bar(stage), phase);
            wait_full += clock64() - t0;
          } else {
            ptx::mbar_wait(full_bar(stage), phase);
          }
          ptx::tcgen05_fence_after_sync();
          const uint64_t adesc = ptx::make_smem_desc_k_sw128(smem_a0 + stage * G::A_STAGE_BYTES);
          const uint64_t bdesc = BMN ? ptx::make_smem_desc_mn_sw128(smem_b0 + stage * G::B_STAGE_BYTES, MN_ATOM_BYTES)
                                     : ptx::make_smem_desc_k_sw128(smem_b0 + stage * G::B_STAGE_BYTES);
          // K-major: advancing K inside the swizzle atom = advancing the start address by 32 B;
          // MN-major: one UMMA_K is UMMA_K_ELEMS k-rows of 128 B.
          constexpr uint32_t UMMA_K_ELEMS = UMMA_K_BYTES / ELEM_BYTES;
          constexpr uint32_t b_step = BMN ? (UMMA_K_ELEMS * 128u) >> 4 : uint32_t(UMMA_K_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K_BYTES / UMMA_K_BYTES; ++k) {
            ptx::umma<KIND, CG>(tmem_d, adesc + uint64_t(k * (UMMA_K_BYTES >> 4)), bdesc + uint64_t(k * b_step),
                                idesc, (kb | uint32_t(k)) != 0u ? 1u : 0u);
          }
          // smem stage reusable / accumulator complete once these MMAs retire
          if (CG == 1) {
            ptx::umma_commit(empty_bar(stage));
            if (kb == num_kb - 1) ptx::umma_commit(tmem_full_bar(as));
          } else {
            ptx::umma_commit_2sm(empty_bar(stage), 0x3);
            if (kb == num_kb - 1) ptx::umma_commit_2sm(tmem_full_bar(as), 0x3);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) {
        dbg[blockIdx.x * 8 + 2] = wait_full;
        dbg[blockIdx.x * 8 + 3] = wait_tmem;
        dbg[blockIdx.x * 8 + 4] = clock64() - t_begin;
      }
    }
  } else if (FUSE_A && warp >= NUM_THREADS / 32) {
    // ================= A rounding (warps 6..9 of every CTA) =================
    const uint32_t pt = threadIdx.x - NUM_THREADS;             // 0..127 within the CTA's prep group
    const uint32_t k4 = fuse.k_elems / 4;
    const uint32_t granules = (rows + PREP_ROWS - 1) / PREP_ROWS;
    // The first raster group (what the first wave of tiles needs) is rounded at full speed by all
    // four warps; after that the GEMM consumes 2048 rows per ~1.2 ms, so only `late_warps` warps
    // keep sweeping: less HBM / L2 pressure on the GEMM's own loads.  Streaming (evict-first)
    // accesses keep the sweep from displacing the A / B^T panels the co-running tiles share in L2.
    const uint32_t first_granules = RASTER_GROUP_ROWS / PREP_ROWS;
    for (uint32_t gr = fuse.pre_done; gr < granules; ++gr) {
      const uint32_t active = (gr < first_granules) ? uint32_t(PREP_WARPS) : fuse.late_warps;
      const size_t begin = size_t(gr) * PREP_ROWS * k4;
      const size_t end = size_t(min(rows, (gr + 1) * PREP_ROWS)) * k4;
      if (pt < active * 32) {
        const size_t stride = size_t(gridDim.x) * (active * 32);
        size_t i = begin + size_t(blockIdx.x) * (active * 32) + pt;
        for (; i + 7 * stride < end; i += 8 * stride) {  // 8 independent 16-byte loads in flight per thread
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcs(fuse.a_raw + i + u * stride);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u].x = round_tf32(v[u].x); v[u].y = round_tf32(v[u].y);
            v[u].z = round_tf32(v[u].z); v[u].w = round_tf32(v[u].w);
            __stcs(fuse.a_prep + i + u * stride, v[u]);
          }
        }
        for (; i < end; i += stride) {
          float4 v = __ldcs(fuse.a_raw + i);
          v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
          __stcs(fuse.a_prep + i, v);
        }
        __threadfence();                                        // this thread's stores visible GPU-wide
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PREP_WARPS * 32) : "memory");  // the CTA's prep warps only
      if (pt == 0) atomicAdd(fuse.a_done + gr, 1u);
    }
  } else {
    // ================= epilogue (warps 2..5 of every CTA) =================
    const uint32_t quarter = warp & 3u;  // TMEM lanes [32*quarter, +32) are this warp's
    uint32_t iter = 0;
    for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++iter) {
      const TileCoord tc = tile_coord(t, tiles_r, tiles_c, raster_group);
      const uint32_t as = iter & 1u;
      const uint32_t aphase = (iter >> 1) & 1u;
      ptx::mbar_wait(tmem_full_bar(as), aphase);
      ptx::tcgen05_fence_after_sync();
      const uint32_t row = tc.r * G::TILE_ROWS + cta_rank * BLOCK_M + quarter * 32 + lane;
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
      if (lane == 0) {
        if (CG == 1) ptx::mbar_arrive(tmem_empty_bar(as));
        else ptx::mbar_arrive_cluster(tmem_empty_bar(as), 0);  // the leader's MMA issuer waits on it
      }
    }
  }

  ptx::tcgen05_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after_sync();
    ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

// ---- operand preparation ------------------------------------------------------------------------

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

// ---- 3xTF32 operand construction (MM_FLAG_TF32X3) ------------------------------------------------
// x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi).  A*B ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b
// (the dropped lo*lo term is 2^-22 relative).  The three products are folded into ONE GEMM with
// K' = 3K by interleaving 16-element k-blocks:  A' = [hi | hi | lo],  B'^T = [hi | lo | hi],
// so the unchanged tcgen05 kernel accumulates all three in its FP32 TMEM accumulator.
constexpr int SPLIT_BLOCK = 16;  // K % 16 == 0 by the reference's shape rule for float

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

// dst[r][3K]: per 16-block of k -> [hi16 | hi16 | lo16]; one thread per float4 of the source row.
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t rows, uint32_t k) {
  const size_t k4 = k / 4;
  const size_t total = rows * k4;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / k4;
    const uint32_t c4 = uint32_t(i - r * k4);       // float4 index within the row
    const uint32_t blk = c4 / (SPLIT_BLOCK / 4), in = c4 % (SPLIT_BLOCK / 4);
    const float4 v = src[i];
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    float4 *row = dst + r * (3 * k4) + size_t(blk) * (3 * SPLIT_BLOCK / 4) + in;
    row[0] = hi;
    row[SPLIT_BLOCK / 4] = hi;
    row[2 * SPLIT_BLOCK / 4] = lo;
  }
}

// src (src_rows = K) x (src_cols) row-major -> dst[c][3K] with per-16-block [a | b | c] where
// B_ORDER selects (hi, lo, hi) for the B operand and (hi, hi, lo) for a transposed A.
template <bool B_ORDER>
__global__ void __launch_bounds__(256)
split3_transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, uint32_t src_rows,
                        uint32_t src_cols) {
  constexpr int TILE = 64;
  __shared__ float tile[TILE][TILE + 1];
  const uint32_t c0 = blockIdx.x * TILE;
  const uint32_t r0 = blockIdx.y * TILE;
  const int x = threadIdx.x % TILE;
  const int y = threadIdx.x / TILE;
#pragma unroll 4
  for (int i = y; i < TILE; i += 4) {
    const uint32_t r = r0 + i, c = c0 + x;
    if (r < src_rows && c < src_cols) tile[i][x] = src[size_t(r) * src_cols + c];
  }
  __syncthreads();
#pragma unroll 4
  for (int i = y; i < TILE; i += 4) {
    const uint32_t c = c0 + i, r = r0 + x;  // dst row = src col; r = k index
    if (c < src_cols && r < src_rows) {
      float hi, lo;
      split_tf32(tile[x][i], hi, lo);
      float *out = dst + size_t(c) * (3 * size_t(src_rows)) + size_t(r / SPLIT_BLOCK) * (3 * SPLIT_BLOCK) +
                   (r % SPLIT_BLOCK);
      out[0] = hi;
      out[SPLIT_BLOCK] = B_ORDER ? lo : hi;
      out[2 * SPLIT_BLOCK] = B_ORDER ? hi : lo;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
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

// MN-major B operand read from row-major B (K x M): box = {64 elements of M (128 B), 64 k-rows}.
int make_b_mn_map(CUtensorMap *map, const void *base, uint64_t k, uint64_t m) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {m, k};
  cuuint64_t gstride[1] = {m * 2};
  cuuint32_t box[2] = {uint32_t(MN_ATOM), uint32_t(BLOCK_K_BYTES / 2)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled (MN-major B) failed with CUresult " + std::to_string(int(r)));
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

// Tail of the scratch: [granule counters of the fused A rounding, 64 KiB][soft wave-barrier counter, 256 B]
constexpr size_t TILE_SYNC_BYTES = 256;
constexpr size_t A_DONE_BYTES = 64 * 1024;  // 16384 granules of 256 rows = 4 Mi rows
constexpr size_t TAIL_BYTES = TILE_SYNC_BYTES + A_DONE_BYTES;

static bool split3(int dtype, int flags) { return dtype == MM_DTYPE_FLOAT && (flags & MM_FLAG_TF32X3); }

// half: B is consumed MN-major straight from the caller's row-major K x M array (no B^T copy).
// MM_TCGEN05_B_MN=0 restores the transposed-copy path for A/B measurements.
bool tcgen05_b_direct(int dtype) {
  static const bool enabled = [] {
    const char *e = std::getenv("MM_TCGEN05_B_MN");
    return !(e && e[0] == '0');
  }();
  return dtype == MM_DTYPE_HALF && enabled;
}

size_t tcgen05_bt_bytes(int dtype, unsigned k, unsigned m, int flags) {
  if (tcgen05_b_direct(dtype)) return 0;
  const size_t eb = (dtype == MM_DTYPE_FLOAT) ? 4 : 2;
  return align_up(size_t(m) * k * eb * (split3(dtype, flags) ? 3 : 1), 1024);
}

size_t tcgen05_scratch_bytes(int dtype, unsigned n, unsigned k, unsigned m, int flags) {
  const size_t eb = (dtype == MM_DTYPE_FLOAT) ? 4 : 2;
  size_t bytes = TAIL_BYTES + tcgen05_bt_bytes(dtype, k, m, flags);  // counters (tail) + B^T
  if (dtype == MM_DTYPE_FLOAT || (flags & MM_FLAG_TRANSPOSED_A)) {
    bytes += align_up(size_t(n) * k * eb * (split3(dtype, flags) ? 3 : 1), 1024);
  }
  return bytes;
}

// Experiment hook (scripts/exp_tf32_rounding.py): feed raw fp32 bits to kind::tf32 to MEASURE the
// truncation bias that motivates the rounding pass.  Never set in production.
static bool experiment_no_round() {
  static const bool v = std::getenv("MM_EXPERIMENT_TF32_NO_ROUND") != nullptr;
  return v;
}

// B (row-major K x M) -> B^T (M x K, K-major MMA operand) into `bt`; float is rounded to TF32.
int tcgen05_prepare_b(int dtype, const void *b, void *bt, unsigned k, unsigned m, int flags,
                      const void **b_op, cudaStream_t stream) {
  *b_op = bt;
  if (tcgen05_b_direct(dtype)) {
    *b_op = b;  // nothing to prepare
    return MM_OK;
  }
  if (split3(dtype, flags)) {
    dim3 grid((m + 63) / 64, (k + 63) / 64);
    split3_transpose_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const float *>(b),
                                                           static_cast<float *>(bt), k, m);
  } else if (dtype == MM_DTYPE_FLOAT) {
    if (experiment_no_round()) {
      launch_transpose<float, false>(b, bt, k, m, stream);
    } else {
      launch_transpose<float, true>(b, bt, k, m, stream);
    }
  } else {
    launch_transpose<__half, false>(b, bt, k, m, stream);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

// `rows` rows of A -> the K-major A operand.  Row-major float A is rounded into `aprep`; row-major
// half A is used in place; A stored K x N (`transposed`, leading dimension n_total, only whole
// matrices) is transposed into `aprep`.  *a_op receives the operand pointer.
int tcgen05_prepare_a(int dtype, const void *a, void *aprep, unsigned rows, unsigned k, int flags,
                      const void **a_op, const void **a_raw, cudaStream_t stream) {
  const bool transposed = (flags & MM_FLAG_TRANSPOSED_A) != 0;
  *a_op = a;
  *a_raw = nullptr;
  if (tcgen05_fuse_a(dtype, flags) && size_t(rows) / PREP_ROWS < A_DONE_BYTES / sizeof(unsigned int)) {
    // The GEMM kernel rounds A into `aprep` itself, in the background.  The rows of the first
    // raster group — what the first wave of tiles needs before it can start — are rounded here by
    // the stand-alone kernel (25 us at 16384^3), so that the GEMM does not begin with a stall.
    const unsigned head = std::min<unsigned>(rows, RASTER_GROUP_ROWS);
    const size_t count4 = size_t(head) * k / 4;
    const int blocks = int(std::min<size_t>((count4 + 255) / 256, size_t(num_sms()) * 16));
    round_tf32_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float4 *>(a), static_cast<float4 *>(aprep),
                                                 count4);
    MM_CUDA_TRY(cudaGetLastError());
    *a_op = aprep;
    *a_raw = a;
    return MM_OK;
  }
  if (split3(dtype, flags)) {
    if (transposed) {
      dim3 grid((rows + 63) / 64, (k + 63) / 64);
      split3_transpose_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const float *>(a),
                                                              static_cast<float *>(aprep), k, rows);
    } else {
      const size_t total4 = size_t(rows) * k / 4;
      const int blocks = int(std::min<size_t>((total4 + 255) / 256, size_t(num_sms()) * 16));
      split3_rows_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float4 *>(a),
                                                    static_cast<float4 *>(aprep), rows, k);
    }
    *a_op = aprep;
  } else if (dtype == MM_DTYPE_FLOAT) {
    if (transposed) {
      if (experiment_no_round()) {
        launch_transpose<float, false>(a, aprep, k, rows, stream);
      } else {
        launch_transpose<float, true>(a, aprep, k, rows, stream);  // A stored K x N -> N x K
      }
      *a_op = aprep;
    } else if (!experiment_no_round()) {
      const size_t count4 = size_t(rows) * k / 4;
      const int blocks = int(std::min<size_t>((count4 + 255) / 256, size_t(num_sms()) * 16));
      round_tf32_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float4 *>(a),
                                                   static_cast<float4 *>(aprep), count4);
      *a_op = aprep;
    }
  } else if (transposed) {
    launch_transpose<__half, false>(a, aprep, k, rows, stream);
    *a_op = aprep;
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

// 1 = single-CTA tiles, 2 = cta_group::2 CTA pairs (default).  MM_TCGEN05_CTA_GROUP overrides for A/B runs.
static int cta_group_choice() {
  static const int v = [] {
    const char *e = std::getenv("MM_TCGEN05_CTA_GROUP");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  return v;
}

template <int KIND, typename TOut, int CG, bool BMN, bool FUSE_A>
static int launch_gemm_variant(const CUtensorMap &map_a, const CUtensorMap &map_b, void *c, unsigned rows,
                               unsigned m, uint32_t k_bytes, unsigned int *tile_sync, FuseA fuse,
                               cudaStream_t stream) {
  using G = Geo<CG>;
  auto kern = gemm_tcgen05_kernel<KIND, TOut, CG, BMN, FUSE_A>;
  MM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G::SMEM_BYTES)));
  const uint32_t tiles = ceil_div(rows, G::TILE_ROWS) * ceil_div(m, BLOCK_N);
  const uint32_t groups = std::min<uint32_t>(tiles, uint32_t(num_sms()) / CG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(NUM_THREADS + (FUSE_A ? PREP_WARPS * 32 : 0));
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // MM_TCGEN05_DEBUG=1 (diagnostics only): per-CTA stall cycle counters, printed after a sync
  static const bool debug = std::getenv("MM_TCGEN05_DEBUG") != nullptr;
  unsigned long long *dbg = nullptr;
  if (debug) {
    MM_CUDA_TRY(cudaMalloc(&dbg, sizeof(unsigned long long) * 8 * cfg.gridDim.x));
    MM_CUDA_TRY(cudaMemsetAsync(dbg, 0, sizeof(unsigned long long) * 8 * cfg.gridDim.x, stream));
  }
  // tuning knobs (diagnostics): ring depth and the TMA loads' L2 eviction priority
  // Ring depth actually used (<= G::STAGES, the allocation).  Measured (profiles/r01_exp_tile_sync.log,
  // float 16384^3 / half 32768^3, CTA pairs, WITH the soft wave barrier): depth 4 leaves the tensor
  // pipe 79-86 % active, depth 5-6 reach 96-98 % at unchanged DRAM traffic (17.7 / 68.8 GB).  Without
  // the barrier depth >= 5 multiplied the DRAM re-reads (27.6 / 31.7 GB at depth 5 / 6, 246 GB for
  // half at depth 6) and lowered the sustained, power-capped rate.
  static const uint32_t stages = [] {
    const char *e = std::getenv("MM_TCGEN05_STAGES");
    const int dflt = G::STAGES;
    const int v = e ? std::atoi(e) : dflt;
    return uint32_t(std::min(std::max(v, 2), int(G::STAGES)));
  }();
  static const uint64_t l2_policy = [] {
    const char *e = std::getenv("MM_TCGEN05_L2");
    if (e && e[0] == 'f') return ptx::L2_EVICT_FIRST;
    if (e && e[0] == 'n') return ptx::L2_EVICT_NORMAL;
    if (e && e[0] == 'l') return ptx::L2_EVICT_LAST;
    return ptx::L2_EVICT_NORMAL;
  }();
  // rows of C per rasterisation group = the height of the patch that co-running tiles share
  // through L2 (the "memory tile" of the reference's I/O model); scripts/tile_sweep.py sweeps it
  static const uint32_t raster_group = [] {
    const char *e = std::getenv("MM_TCGEN05_RASTER_ROWS");
    const int rows_per_group = e ? std::atoi(e) : RASTER_GROUP_ROWS;
    return uint32_t(std::max(1, rows_per_group / G::TILE_ROWS));
  }();
  static const bool use_tile_sync = [] {
    const char *e = std::getenv("MM_TCGEN05_TILE_SYNC");
    return !(e && e[0] == '0');
  }();
  if (!use_tile_sync) tile_sync = nullptr;
  if (tile_sync) MM_CUDA_TRY(cudaMemsetAsync(tile_sync, 0, sizeof(unsigned int), stream));
  if (FUSE_A) {
    const size_t granules = (size_t(rows) + PREP_ROWS - 1) / PREP_ROWS;
    MM_CUDA_TRY(cudaMemsetAsync(fuse.a_done, 0, granules * sizeof(unsigned int), stream));
  }
  MM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, static_cast<TOut *>(c), uint32_t(rows), uint32_t(m),
                                 k_bytes, stages, raster_group, l2_policy, tile_sync, fuse, dbg));
  if (debug) {
    MM_CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<unsigned long long> h(8 * cfg.gridDim.x);
    MM_CUDA_TRY(cudaMemcpy(h.data(), dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    double s[5] = {0, 0, 0, 0, 0};
    int nlead = 0;
    for (unsigned b = 0; b < cfg.gridDim.x; ++b) {
      s[0] += double(h[b * 8 + 0]);
      s[1] += double(h[b * 8 + 1]);
      if (h[b * 8 + 4]) {
        ++nlead;
        s[2] += double(h[b * 8 + 2]);
        s[3] += double(h[b * 8 + 3]);
        s[4] += double(h[b * 8 + 4]);
      }
    }
    fprintf(stderr, "[tcgen05 debug CG=%d rows=%u m=%u] producer: wait_empty %.1f%% of %.0f cyc | mma: wait_full %.1f%% wait_tmem %.1f%% of %.0f cyc\n",
            CG, rows, m, 100.0 * s[0] / s[1], s[1] / cfg.gridDim.x, 100.0 * s[2] / s[4], 100.0 * s[3] / s[4], s[4] / nlead);
  }
  return MM_OK;
}

// float, row-major A, single-pass TF32: A's rounding can be fused into the GEMM kernel (FuseA).
// EXPERIMENTAL, off by default (MM_TCGEN05_FUSE_A=1 enables it): measured over three variants
// (profiles/r01_exp_fuse_a*.log) the background sweep slows the power- and bandwidth-sharing GEMM by
// about what the separate 0.34 ms pass costs (step 10.47-10.68 ms fused vs 10.37-10.80 ms separate),
// so the simpler path without a cross-CTA wait stays the default.
bool tcgen05_fuse_a(int dtype, int flags) {
  static const bool enabled = [] {
    const char *e = std::getenv("MM_TCGEN05_FUSE_A");
    return e && e[0] == '1';
  }();
  return enabled && dtype == MM_DTYPE_FLOAT && !(flags & (MM_FLAG_TRANSPOSED_A | MM_FLAG_TF32X3)) &&
         !experiment_no_round();
}

// C[rows x m] = Aop[rows x k] * B on the tensor cores; `b_op` is the K-major copy B^T (m x k), or
// the caller's row-major B (k x m) when tcgen05_b_direct(dtype).  `a_raw` non-null: a_op is the
// (not yet written) rounded-copy buffer and the kernel rounds `a_raw` into it itself; `counters`
// then provides the per-granule completion counters.
int tcgen05_gemm(int dtype, const void *a_op, const void *b_op, void *c, unsigned rows, unsigned k,
                 unsigned m, int flags, unsigned int *tile_sync, const void *a_raw, unsigned int *counters,
                 cudaStream_t stream) {
  const unsigned k_orig = k;
  if (split3(dtype, flags)) k *= 3;  // the operands carry [hi|hi|lo] x [hi|lo|hi] per 16-block of K
  const bool is_f32 = dtype == MM_DTYPE_FLOAT;
  const size_t eb = is_f32 ? 4 : 2;
  const int cg = cta_group_choice();
  const bool bmn = tcgen05_b_direct(dtype);
  CUtensorMap map_a, map_b;
  int rc = make_operand_map(&map_a, a_op, dtype, rows, k, BLOCK_M);
  if (rc != MM_OK) return rc;
  rc = bmn ? make_b_mn_map(&map_b, b_op, k, m)
           : make_operand_map(&map_b, b_op, dtype, m, k, cg == 2 ? Geo<2>::LOAD_N : Geo<1>::LOAD_N);
  if (rc != MM_OK) return rc;
  const uint32_t k_bytes = uint32_t(size_t(k) * eb);
  static const uint32_t late_warps = [] {
    const char *e = std::getenv("MM_TCGEN05_FUSE_A_LATE_WARPS");
    return uint32_t(std::min(std::max(e ? std::atoi(e) : 1, 1), PREP_WARPS));
  }();
  // leading granules covered by tcgen05_prepare_a's stand-alone pass over the first raster group
  const unsigned head = std::min<unsigned>(rows, RASTER_GROUP_ROWS);
  const uint32_t pre_done = (head == rows) ? (rows + PREP_ROWS - 1) / PREP_ROWS : head / PREP_ROWS;
  FuseA fuse{static_cast<const float4 *>(a_raw), static_cast<float4 *>(const_cast<void *>(a_op)), counters, k_orig,
             late_warps, pre_done};
  if (is_f32 && a_raw != nullptr) {
    return cg == 2 ? launch_gemm_variant<ptx::KIND_TF32, float, 2, false, true>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream)
                   : launch_gemm_variant<ptx::KIND_TF32, float, 1, false, true>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream);
  }
  if (is_f32) {
    return cg == 2 ? launch_gemm_variant<ptx::KIND_TF32, float, 2, false, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream)
                   : launch_gemm_variant<ptx::KIND_TF32, float, 1, false, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream);
  }
  if (bmn) {
    return cg == 2 ? launch_gemm_variant<ptx::KIND_F16, __half, 2, true, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream)
                   : launch_gemm_variant<ptx::KIND_F16, __half, 1, true, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream);
  }
  return cg == 2 ? launch_gemm_variant<ptx::KIND_F16, __half, 2, false, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream)
                 : launch_gemm_variant<ptx::KIND_F16, __half, 1, false, false>(map_a, map_b, c, rows, m, k_bytes, tile_sync, fuse, stream);
}

int launch_tcgen05(int dtype, const GemmArgs &g, void *scratch, size_t scratch_bytes) {
  if (dtype != MM_DTYPE_FLOAT && dtype != MM_DTYPE_HALF) {
    return fail(MM_ERR_UNSUPPORTED, "tcgen05 path handles float and half only");
  }
  if (scratch_bytes < tcgen05_scratch_bytes(dtype, g.n, g.k, g.m, g.flags)) {
    return fail(MM_ERR_INVALID, "tcgen05 scratch too small");
  }
  if (g.dry_run) {
    // force the lazily loaded kernels in (prep + GEMM variants) and the driver entry point
    cudaFuncAttributes attr;
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, round_tf32_kernel));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, transpose_prep_kernel<float, true>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, transpose_prep_kernel<__half, false>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, gemm_tcgen05_kernel<ptx::KIND_TF32, float, 2, false, true>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, gemm_tcgen05_kernel<ptx::KIND_TF32, float, 2, false, false>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, gemm_tcgen05_kernel<ptx::KIND_F16, __half, 2, true, false>));
    if (!get_encode_fn()) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return MM_OK;
  }
  unsigned char *sp = static_cast<unsigned char *>(scratch);
  void *bt = sp;
  void *aprep = sp + tcgen05_bt_bytes(dtype, g.k, g.m, g.flags);
  unsigned int *tile_sync = reinterpret_cast<unsigned int *>(sp + scratch_bytes - TILE_SYNC_BYTES);
  unsigned int *a_done = reinterpret_cast<unsigned int *>(sp + scratch_bytes - TAIL_BYTES);

  const void *b_op = nullptr;
  int rc = tcgen05_prepare_b(dtype, g.b, bt, g.k, g.m, g.flags, &b_op, g.stream);
  if (rc != MM_OK) return rc;
  const void *a_op = nullptr, *a_raw = nullptr;
  rc = tcgen05_prepare_a(dtype, g.a, aprep, g.n, g.k, g.flags, &a_op, &a_raw, g.stream);
  if (rc != MM_OK) return rc;
  if (g.ev_prep_done) MM_CUDA_TRY(cudaEventRecord(g.ev_prep_done, g.stream));
  return tcgen05_gemm(dtype, a_op, b_op, g.c, g.n, g.k, g.m, g.flags, tile_sync, a_raw, a_done, g.stream);
}

}  // namespace mm
