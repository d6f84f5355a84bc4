// Tensor-core path of the hot path for the dense (Multiply, Add) contraction on float, half and uint8_t:
//   C[N x M] = A[N x K] * B[K x M]
// B200 counterpart of the reference's PE chain + streamers (kernel/Compute.cpp:53-146,
// kernel/Memory.cpp:58-438) for MM_MAP_OP=Multiply, MM_REDUCE_OP=Add, MM_DATA_TYPE in {float, half, uint8_t}
// (uint8_t: kind::i8 with exact 32-bit integer accumulation, truncated to 8 bits in the epilogue = the reference's
// arithmetic modulo 256, bit for bit; reference CMakeLists.txt:43-46).
//
// Structure (one persistent CTA per SM, warp-specialised, no CUTLASS):
//   warp 0   TMA producer (the role of ReadA / ReadB / FeedB): cp.async.bulk.tensor of 128-byte-swizzled
//            A (128 x BK, K-major) and B tiles into a STAGES-deep shared-memory ring, mbarrier full/empty
//            pairs.  B is read MN-major STRAIGHT from the reference's row-major K x M layout (boxes of
//            one swizzle atom of columns x BK k-rows), or K-major from a transposed copy (tuning knob).
//   warp 1   MMA issuer (the PE chain): one thread issues tcgen05.mma (kind::tf32 | kind::f16, M = 128 per
//            CTA, N = 128 | 256, K = 32 bytes) into one of two FP32 accumulators in TMEM; tcgen05.commit
//            releases smem stages and publishes finished accumulators
//   warps 2-5 epilogue (WriteC, kernel/Memory.cpp:361-392): tcgen05.ld the accumulator (each warp its
//            32-lane TMEM quarter), convert, stage 32 x 32 blocks in swizzled shared memory and write
//            them with TMA stores (cp.async.bulk.tensor, clipped to n < N, m < M by the tensor map);
//            overlaps the next tile's main loop (double-buffered TMEM)
//
// Operand preparation (O(N*K + K*M) bytes against O(N*K*M) flops):
//   * kind::tf32 reads only the upper 19 bits of each fp32 operand, i.e. it TRUNCATES.  The
//     reference's inputs are all positive (U[1,10], test/TestSimulation.cpp:46-55), so truncation
//     would bias every product by about -2^-11 * 2 and land the sum right at the 1e-3 tolerance.
//     A and B are therefore rounded to nearest TF32 (cvt.rna.tf32.f32) into scratch copies first.
//   * B's preparation CAN run concurrently with the GEMM (tuning knob b_overlap): a co-resident
//     persistent kernel on a second stream prepares B panel by panel (BLOCK_N columns x all of K, in
//     the order the rasterisation consumes panels) and publishes each panel through a counter; the
//     GEMM's producer waits for a panel's counter before its first TMA load from it.  For the LOCAL
//     rounding pass this is off by default: measured (profiles/r02_exp_b_overlap.md) the HBM-bound pass
//     slows the co-running, latency-sensitive GEMM by more than it hides (1.83 vs 1.64 ms per step on a
//     2048-row block, 11.38 vs 10.97 ms at 16384^3).  The same pass (overlapped or not) reads row-slices
//     of B from PEER GPUs in the multi-GPU host path: the NVLink all-gather of B is fused into it.
//   * half needs no preparation at all: A is K-major as stored, B is read MN-major in place.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_host.cuh"

namespace mm {
namespace {

constexpr int BLOCK_M = 128;          // C rows per CTA      (UMMA M = 128 * CTA group size)
constexpr int BLOCK_K_BYTES = 128;    // one 128-byte swizzle atom of K per stage
constexpr int UMMA_K_BYTES = 32;      // K extent of one tcgen05.mma
constexpr int ACC_STAGES = 2;
constexpr int NUM_THREADS = 192;
constexpr int EPI_BUF_BYTES = 4096;   // one 32 x 32 fp32 block per epilogue warp
constexpr int EPI_BYTES = 4 * EPI_BUF_BYTES;
constexpr int BAR_BYTES = 256;
constexpr int MAX_DYN_SMEM = 232448;  // 227 KiB per CTA on sm_100a

// Per-variant geometry.  CG = 1: one CTA computes a 128 x BN tile and stages A (128 rows) + all BN
// columns of B per k-block.  CG = 2 (cta_group::2): a CTA PAIR computes 256 x BN with ONE tcgen05.mma
// per k-step issued by the leader CTA; each CTA stages only its own 128 A rows and its half of the B
// tile, so per-SM shared-memory fill and L2 traffic drop and the freed memory deepens the ring.
template <int CG, int BN>
struct Geo {
  static constexpr int LOAD_N = BN / CG;                            // B columns staged per CTA
  static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K_BYTES;     // 16 KiB
  static constexpr int B_STAGE_BYTES = LOAD_N * BLOCK_K_BYTES;      // 8 .. 32 KiB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int FIT = (MAX_DYN_SMEM - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int MAX_STAGES = FIT < 8 ? FIT : 8;              // 6 (2,256)  4 (1,256)  8 (2,128)  6 (1,128)
  static constexpr int TILE_ROWS = BLOCK_M * CG;                    // C rows per CTA group
  static constexpr int TMEM_COLS = ACC_STAGES * BN;                 // 512 | 256
  static constexpr size_t smem_bytes(int stages) {
    return size_t(stages) * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + BAR_BYTES;
  }
};

struct TileCoord {
  uint32_t r, c;
};

// Grouped rasterisation: RASTER_GROUP row-tiles sweep all column-tiles together so that the
// concurrently running tiles share A row-panels and B column-panels through L2.  Column tiles are
// visited in ascending order within a group — the order B's preparation publishes its panels in.
__device__ __forceinline__ TileCoord tile_coord(uint32_t t, uint32_t tiles_r, uint32_t tiles_c,
                                                uint32_t raster_group) {
  const uint32_t per_group = raster_group * tiles_c;
  const uint32_t g = t / per_group;
  const uint32_t first = g * raster_group;
  const uint32_t gsize = min(raster_group, tiles_r - first);
  const uint32_t in = t - g * per_group;
  return TileCoord{first + in % gsize, in / gsize};
}

// ---- epilogue stores ------------------------------------------------------------------------------
// Direct variant (tuning knob tma_store = 0): each lane owns one C row of the chunk.
template <typename TOut>
__device__ __forceinline__ void store_chunk(TOut *crow, const uint32_t (&v)[32], uint32_t col, uint32_t cols);

template <>
__device__ __forceinline__ void store_chunk<float>(float *crow, const uint32_t (&v)[32], uint32_t col,
                                                   uint32_t cols) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if (col + j + 4 <= cols) {
      *reinterpret_cast<uint4 *>(crow + col + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}

__device__ __forceinline__ void pack_half(const uint32_t (&v)[32], uint32_t (&p)[16]) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    __half2 h = __floats2half2_rn(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
    p[q] = *reinterpret_cast<uint32_t *>(&h);
  }
}

template <>
__device__ __forceinline__ void store_chunk<__half>(__half *crow, const uint32_t (&v)[32], uint32_t col,
                                                    uint32_t cols) {
  uint32_t p[16];
  pack_half(v, p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (col + 8 * j + 8 <= cols) {
      *reinterpret_cast<uint4 *>(crow + col + 8 * j) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
    }
  }
}

// uint8_t: the accumulator is the exact 32-bit sum; its low byte is the reference's result (arithmetic modulo 256).
__device__ __forceinline__ void pack_u8(const uint32_t (&v)[32], uint32_t (&p)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    p[q] = (v[4 * q] & 0xFFu) | ((v[4 * q + 1] & 0xFFu) << 8) | ((v[4 * q + 2] & 0xFFu) << 16) | (v[4 * q + 3] << 24);
  }
}

template <>
__device__ __forceinline__ void store_chunk<unsigned char>(unsigned char *crow, const uint32_t (&v)[32], uint32_t col,
                                                           uint32_t cols) {
  uint32_t p[8];
  pack_u8(v, p);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (col + 16 * j + 16 <= cols) {
      *reinterpret_cast<uint4 *>(crow + col + 16 * j) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
    }
  }
}

// Staged variant: lane = row of a 32 x 32 block; the block is written into shared memory in the
// swizzled layout the C tensor map expects (row pitch 128 B with SWIZZLE_128B for float, 64 B with
// SWIZZLE_64B for half: 16-byte chunk index XOR row bits), so the quarter-warp phases of the 128-bit
// shared stores hit distinct banks, and ONE TMA store writes the whole block as full 128-byte lines.
template <typename TOut>
__device__ __forceinline__ void stage_chunk(uint32_t buf, uint32_t lane, const uint32_t (&v)[32]);

template <>
__device__ __forceinline__ void stage_chunk<float>(uint32_t buf, uint32_t lane, const uint32_t (&v)[32]) {
  const uint32_t row = buf + lane * 128u;
#pragma unroll
  for (uint32_t j = 0; j < 8; ++j) {
    ptx::st_shared_v4(row + ((j ^ (lane & 7u)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

template <>
__device__ __forceinline__ void stage_chunk<__half>(uint32_t buf, uint32_t lane, const uint32_t (&v)[32]) {
  uint32_t p[16];
  pack_half(v, p);
  const uint32_t row = buf + lane * 64u;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    ptx::st_shared_v4(row + ((j ^ ((lane >> 1) & 3u)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
  }
}

template <>
__device__ __forceinline__ void stage_chunk<unsigned char>(uint32_t buf, uint32_t lane, const uint32_t (&v)[32]) {
  uint32_t p[8];
  pack_u8(v, p);
  const uint32_t row = buf + lane * 32u;   // 32-byte rows, SWIZZLE_32B: 16-byte chunk index XOR address bit 7
#pragma unroll
  for (uint32_t j = 0; j < 2; ++j) {
    ptx::st_shared_v4(row + ((j ^ ((lane >> 2) & 1u)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
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
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Run-time launch parameters of the GEMM kernel (one struct so that the 16 instantiations share a
// signature).
struct GemmParams {
  uint32_t rows, cols, k_bytes;
  uint32_t num_stages;       // ring depth actually used (<= Geo::MAX_STAGES, what the smem allocation holds)
  uint32_t raster_group;     // row tiles per rasterisation group
  uint32_t tma_store;        // 1: staged TMA-store epilogue, 0: direct per-lane stores
  uint32_t b_ready_target;   // see b_ready
  uint64_t l2_policy;
  unsigned int *tile_sync;        // soft wave-barrier counter or null
  const unsigned int *b_ready;    // per column tile: preparation items finished, or null (B complete)
};

// C[rows x cols] = A'[rows x k] * B ; A' K-major.  B either MN-major straight from a row-major
// k x cols array (BMN) or K-major from a cols x k transposed copy.  CG == 2 must be launched with
// cluster dimension (2, 1, 1).
template <int KIND, typename TOut, int CG, int BN, bool BMN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, TOut *__restrict__ C, const GemmParams p) {
  using G = Geo<CG, BN>;
  constexpr int ELEM_BYTES = (KIND == ptx::KIND_TF32) ? 4 : (KIND == ptx::KIND_I8 ? 1 : 2);
  constexpr int BLOCK_K_ELEMS = BLOCK_K_BYTES / ELEM_BYTES;
  constexpr int MN_ATOM = 128 / ELEM_BYTES;                 // columns per MN-major swizzle atom (128 B)
  constexpr int MN_ATOM_BYTES = BLOCK_K_ELEMS * 128;        // one atom: BLOCK_K k-rows x 128 B
  const int STAGES = int(p.num_stages);
  const uint32_t rows = p.rows, cols = p.cols;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 128B-swizzled tiles must start on a 1024-byte boundary (same offset in both CTAs of a pair).
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + STAGES * G::A_STAGE_BYTES;
  const uint32_t epi0 = smem_base + STAGES * G::STAGE_BYTES;
  const uint32_t bar_base = epi0 + EPI_BYTES;
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
  const uint32_t tiles_c = (cols + BN - 1) / BN;
  const uint32_t num_tiles = tiles_r * tiles_c;
  const uint32_t num_kb = (p.k_bytes + BLOCK_K_BYTES - 1) / BLOCK_K_BYTES;

  if (CG == 2) cluster_sync_all();  // both CTAs resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    if (p.tma_store) ptx::prefetch_tensormap(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      // ONE arrival: the (leader's) producer's arrive.expect_tx, which books the bytes of BOTH CTAs.
      // The peer's TMA transactions complete_tx on the leader's barrier; no peer arrival is needed
      // (a remote mbarrier.arrive per k-block stalled the peer's producer for ~400 cycles and starved
      // the MMA 68 % of the time).  The phase cannot complete early because the expected byte count
      // includes the peer's half, and the peer cannot run a phase ahead because its empty barrier is
      // released by the same tcgen05.commit.
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);  // tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);        // tcgen05.commit
      ptx::mbar_init(tmem_empty_bar(s), 4 * CG);  // one arrive per epilogue warp of the group (leader's copy)
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<CG>(tmem_slot, G::TMEM_COLS);
  }
  ptx::tcgen05_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  ptx::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer (one per CTA) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t tile_iter = 0;
      int32_t ready_panel = -1;
      for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++tile_iter) {
        // Soft wave barrier: do not start fetching tile #j before every CTA group has finished
        // fetching its tile #(j-1).  A ring deep enough to hide DRAM latency removes the L2-miss
        // back-pressure that otherwise keeps the co-running tiles in lock-step, the groups drift,
        // and the A / B panels they share are re-read from DRAM (measured: 17.7 -> 32 GB at
        // 16384^3).  Purely a performance hint: the wait is bounded, correctness never depends on it.
        if (p.tile_sync != nullptr && tile_iter > 0) {
          const uint32_t target = min(tile_iter * num_groups, num_tiles);
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile unsigned int *>(p.tile_sync) < target) {
            if (clock64() - t0 > 100000) break;  // ~50 us: give up, stay correct
          }
        }
        const TileCoord tc = tile_coord(t, tiles_r, tiles_c, p.raster_group);
        const int32_t a_row = tc.r * G::TILE_ROWS + cta_rank * BLOCK_M;
        const int32_t b_col = tc.c * BN + cta_rank * G::LOAD_N;
        if (p.b_ready != nullptr && int32_t(tc.c) != ready_panel) {
          // B's preparation kernel was ENQUEUED before this kernel and needs no resource this kernel
          // holds, so it always makes progress; the bound only turns an impossible wait into a trap.
          const unsigned int *flag = p.b_ready + tc.c;
          const long long t0 = clock64();
          while (ld_acquire_gpu(flag) < p.b_ready_target) {
            if (clock64() - t0 > (1ll << 34)) __trap();
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy writes -> TMA reads
          ready_panel = int32_t(tc.c);
        }
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_a0 + stage * G::A_STAGE_BYTES;
          const uint32_t sb = smem_b0 + stage * G::B_STAGE_BYTES;
          const int32_t k0 = kb * BLOCK_K_ELEMS;
          if (CG == 1) {
            ptx::mbar_arrive_expect_tx(full_bar(stage), G::STAGE_BYTES);
            ptx::tma_load_2d(sa, &tmap_a, full_bar(stage), k0, a_row, p.l2_policy);
            if (BMN) {
#pragma unroll
              for (int j = 0; j < G::LOAD_N / MN_ATOM; ++j) {
                ptx::tma_load_2d(sb + j * MN_ATOM_BYTES, &tmap_b, full_bar(stage), b_col + j * MN_ATOM, k0, p.l2_policy);
              }
            } else {
              ptx::tma_load_2d(sb, &tmap_b, full_bar(stage), k0, b_col, p.l2_policy);
            }
          } else {
            // both CTAs' bytes are accounted on the LEADER's barrier (peer bit 24 cleared)
            const uint32_t leader_bar = full_bar(stage) & 0xFEFFFFFFu;
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * G::STAGE_BYTES);
            ptx::tma_load_2d_2sm(sa, &tmap_a, leader_bar, k0, a_row, p.l2_policy);
            if (BMN) {
#pragma unroll
              for (int j = 0; j < G::LOAD_N / MN_ATOM; ++j) {
                ptx::tma_load_2d_2sm(sb + j * MN_ATOM_BYTES, &tmap_b, leader_bar, b_col + j * MN_ATOM, k0, p.l2_policy);
              }
            } else {
              ptx::tma_load_2d_2sm(sb, &tmap_b, leader_bar, k0, b_col, p.l2_policy);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.tile_sync != nullptr && cta_rank == 0) atomicAdd(p.tile_sync, 1u);  // this group fetched its tile
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(KIND, BLOCK_M * CG, BN, BMN);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++iter) {
        const uint32_t as = iter & 1u;
        const uint32_t aphase = (iter >> 1) & 1u;
        ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        ptx::tcgen05_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after_sync();
          const uint64_t adesc = ptx::make_smem_desc_k_sw128(smem_a0 + stage * G::A_STAGE_BYTES);
          // MN-major B: 16-bit types swizzle 16-byte chunks over 8 k-rows (layout type 2, SBO 1024 B); tf32 has
          // only the 32-byte-chunk mode over 4 k-rows (layout type 1, SBO 512 B) — see ptx_sm100.cuh
          const uint64_t bdesc = BMN ? ptx::make_smem_desc_mn(smem_b0 + stage * G::B_STAGE_BYTES, MN_ATOM_BYTES,
                                                              ELEM_BYTES == 4 ? 512u : 1024u, ELEM_BYTES == 4 ? 1u : 2u)
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
    }
  } else {
    // ================= epilogue (warps 2..5 of every CTA) =================
    const uint32_t quarter = warp & 3u;  // TMEM lanes [32*quarter, +32) are this warp's
    const uint32_t buf = epi0 + quarter * EPI_BUF_BYTES;
    uint32_t iter = 0;
    for (uint32_t t = group_id; t < num_tiles; t += num_groups, ++iter) {
      const TileCoord tc = tile_coord(t, tiles_r, tiles_c, p.raster_group);
      const uint32_t as = iter & 1u;
      const uint32_t aphase = (iter >> 1) & 1u;
      ptx::mbar_wait(tmem_full_bar(as), aphase);
      ptx::tcgen05_fence_after_sync();
      const uint32_t row0 = tc.r * G::TILE_ROWS + cta_rank * BLOCK_M + quarter * 32;
      const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + as * BN;
      if (p.tma_store) {
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr0 + chunk * 32, v);
          ptx::tmem_ld_wait();
          const uint32_t col = tc.c * BN + chunk * 32;
          if (row0 < rows && col < cols) {                        // warp-uniform
            if (lane == 0) ptx::tma_store_wait_read<0>();         // the previous block has left the buffer
            __syncwarp();
            stage_chunk<TOut>(buf, lane, v);
            ptx::fence_proxy_async_smem();                        // generic-proxy smem writes -> TMA read
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tmap_c, buf, int32_t(col), int32_t(row0));  // clipped to rows x cols by the map
              ptx::tma_store_commit();
            }
          }
        }
      } else {
        const uint32_t row = row0 + lane;
        TOut *crow = C + size_t(row) * cols;
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr0 + chunk * 32, v);
          ptx::tmem_ld_wait();
          if (row < rows) store_chunk<TOut>(crow, v, tc.c * BN + chunk * 32, cols);
        }
      }
      ptx::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) ptx::mbar_arrive(tmem_empty_bar(as));
        else ptx::mbar_arrive_cluster(tmem_empty_bar(as), 0);  // the leader's MMA issuer waits on it
      }
    }
    if (p.tma_store && lane == 0) ptx::tma_store_wait_all<0>();  // stores complete before the CTA's smem goes away
  }

  ptx::tcgen05_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after_sync();
    ptx::tmem_dealloc<CG>(tmem_base, G::TMEM_COLS);
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

// B (row-major K x M, 16-byte vectors) -> dst (same layout), panel by panel: a work item is
// PANEL_ROWS k-rows of one panel of `panel_v` vectors per row; items are numbered panel-major and
// dealt round-robin to the CTAs of a persistent grid, so panels complete in ascending order.  Each
// finished item bumps ready[panel] (release pattern: every thread fences its stores, the CTA syncs,
// one thread adds).  ROUND: elements are floats rounded to nearest TF32; otherwise a plain copy.
// `parts` non-null: k-row r is read from parts[r / part_rows] — full-size K x M arrays on (peer) GPUs
// of which only that slice of rows is valid; rows whose source IS the destination are skipped.
constexpr int PREP_THREADS = 512;
constexpr int PREP_WARPS = PREP_THREADS / 32;
constexpr int PANEL_ROWS = 64;
constexpr int PANEL_ROW_SLOTS = PANEL_ROWS / PREP_WARPS;  // rows per warp per item (4)

template <bool ROUND>
__device__ __forceinline__ uint4 prep_vec(uint4 v) {
  if (ROUND) {
    v.x = __float_as_uint(round_tf32(__uint_as_float(v.x)));
    v.y = __float_as_uint(round_tf32(__uint_as_float(v.y)));
    v.z = __float_as_uint(round_tf32(__uint_as_float(v.z)));
    v.w = __float_as_uint(round_tf32(__uint_as_float(v.w)));
  }
  return v;
}

template <bool ROUND>
__global__ void __launch_bounds__(PREP_THREADS)
prep_b_panels_kernel(const uint4 *__restrict__ single, const uint4 *const *__restrict__ parts, uint32_t part_rows,
                     uint4 *__restrict__ dst, uint32_t k, uint32_t row_v, uint32_t panel_v,
                     unsigned int *__restrict__ ready) {
  const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t panels = (row_v + panel_v - 1) / panel_v;
  const uint32_t items_per_panel = (k + PANEL_ROWS - 1) / PANEL_ROWS;
  const uint32_t items = panels * items_per_panel;
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
    const uint32_t panel = item / items_per_panel;
    const uint32_t r0 = (item - panel * items_per_panel) * PANEL_ROWS;
    const uint32_t v0 = panel * panel_v;
    const uint32_t w = min(panel_v, row_v - v0);
    for (uint32_t c0 = 0; c0 < w; c0 += 64) {   // 64 vectors (1 KiB) of a row per pass: 8 loads in flight per thread
      uint4 buf[PANEL_ROW_SLOTS][2];
      bool live[PANEL_ROW_SLOTS][2];
#pragma unroll
      for (int u = 0; u < PANEL_ROW_SLOTS; ++u) {
        const uint32_t r = r0 + warp + u * PREP_WARPS;
        const uint4 *src = single;
        if (parts != nullptr && r < k) src = parts[r / part_rows];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t c = c0 + lane + 32 * h;
          const size_t off = size_t(r) * row_v + v0 + c;
          live[u][h] = (r < k) && (c < w) && (src + off != static_cast<const uint4 *>(dst) + off);
          if (live[u][h]) buf[u][h] = src[off];
        }
      }
#pragma unroll
      for (int u = 0; u < PANEL_ROW_SLOTS; ++u) {
        const uint32_t r = r0 + warp + u * PREP_WARPS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t c = c0 + lane + 32 * h;
          if (live[u][h]) dst[size_t(r) * row_v + v0 + c] = prep_vec<ROUND>(buf[u][h]);
        }
      }
    }
    if (ready != nullptr) {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(ready + panel, 1u);
    }
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
CUtensorMapDataType tma_dtype(int dtype) {
  return dtype == MM_DTYPE_FLOAT ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (dtype == MM_DTYPE_UINT8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}
uint32_t elem_bytes(int dtype) { return dtype == MM_DTYPE_FLOAT ? 4u : (dtype == MM_DTYPE_UINT8 ? 1u : 2u); }

int encode(CUtensorMap *map, CUtensorMapDataType dt, const void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
           uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle, const char *what) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(MM_ERR_CUDA, std::string("cuTensorMapEncodeTiled (") + what + ") failed with CUresult " +
                                 std::to_string(int(r)));
  }
  return MM_OK;
}

// K-major operand: `rows` rows of `k_elems` elements; box = {128 bytes of K, box_rows}, 128-byte
// swizzle, out-of-bounds reads return zeros (neutral for (Multiply, Add) — SURVEY.md section 5 trap 3).
int make_operand_map(CUtensorMap *map, const void *base, int dtype, uint64_t rows, uint64_t k_elems, uint32_t box_rows) {
  const uint32_t eb = elem_bytes(dtype);
  return encode(map, tma_dtype(dtype), base, k_elems, rows, k_elems * eb, uint32_t(BLOCK_K_BYTES / eb), box_rows,
                CU_TENSOR_MAP_SWIZZLE_128B, "K-major operand");
}

// MN-major B operand read from row-major B (K x M): box = {one 128-byte atom of columns, BLOCK_K k-rows}.
int make_b_mn_map(CUtensorMap *map, const void *base, int dtype, uint64_t k, uint64_t m) {
  const uint32_t eb = elem_bytes(dtype);
  // tf32: tcgen05 reads MN-major 32-bit operands only in the 32-byte-chunk swizzle (SWIZZLE_128B_BASE32B)
  return encode(map, tma_dtype(dtype), base, m, k, m * eb, 128 / eb, uint32_t(BLOCK_K_BYTES / eb),
                eb == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, "MN-major B");
}

// C (row-major rows x m) for the epilogue's TMA stores: 32 x 32 blocks, swizzle = row pitch of the block.
int make_c_map(CUtensorMap *map, void *base, int dtype, uint64_t rows, uint64_t m) {
  const uint32_t eb = elem_bytes(dtype);
  return encode(map, tma_dtype(dtype), base, m, rows, m * eb, 32, 32,
                eb == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : (eb == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), "C");
}

int num_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T, bool ROUND>
void launch_transpose(const void *src, void *dst, uint32_t src_rows, uint32_t src_cols, cudaStream_t stream) {
  dim3 grid((src_cols + 63) / 64, (src_rows + 63) / 64);
  transpose_prep_kernel<T, ROUND><<<grid, 256, 0, stream>>>(static_cast<const T *>(src), static_cast<T *>(dst),
                                                           src_rows, src_cols);
}

bool split3(int dtype, int flags) { return dtype == MM_DTYPE_FLOAT && (flags & MM_FLAG_TF32X3); }

int launch_panels(bool round, const BSource &src, void *dst, size_t elem_bytes, unsigned k, unsigned m,
                  unsigned panel_cols, unsigned int *ready, int grid, cudaStream_t stream) {
  const uint32_t row_v = uint32_t(size_t(m) * elem_bytes / 16);
  const uint32_t panel_v = uint32_t(size_t(panel_cols) * elem_bytes / 16);
  const uint4 *single = static_cast<const uint4 *>(src.b);
  const uint4 *const *parts = reinterpret_cast<const uint4 *const *>(src.src);
  if (round) {
    prep_b_panels_kernel<true><<<grid, PREP_THREADS, 0, stream>>>(single, parts, src.part_rows, static_cast<uint4 *>(dst),
                                                                 k, row_v, panel_v, ready);
  } else {
    prep_b_panels_kernel<false><<<grid, PREP_THREADS, 0, stream>>>(single, parts, src.part_rows, static_cast<uint4 *>(dst),
                                                                  k, row_v, panel_v, ready);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

}  // namespace

// Tail of the scratch: [panel counters of B's preparation, 64 KiB][soft wave-barrier counter, 256 B]
constexpr size_t TILE_SYNC_BYTES = 256;
constexpr size_t B_READY_BYTES = 64 * 1024;  // 16384 panels of >= 128 columns
constexpr size_t TAIL_BYTES = TILE_SYNC_BYTES + B_READY_BYTES;
static_assert(TAIL_BYTES == kTcgen05TailBytes, "common.cuh and gemm_tcgen05.cu disagree on the scratch tail");

Tcgen05Counters tcgen05_counters(void *scratch, size_t scratch_bytes) {
  unsigned char *tail = static_cast<unsigned char *>(scratch) + scratch_bytes;
  return Tcgen05Counters{reinterpret_cast<unsigned int *>(tail - TILE_SYNC_BYTES),
                         reinterpret_cast<unsigned int *>(tail - TAIL_BYTES)};
}

bool tcgen05_b_mn(int dtype, int flags, const Tuning &t) { return t.b_mn() && !split3(dtype, flags); }

bool tcgen05_b_in_place(int dtype, int flags, const Tuning &t) {
  return tcgen05_b_mn(dtype, flags, t) && (dtype == MM_DTYPE_HALF || dtype == MM_DTYPE_UINT8 || t.tf32_no_round());
}

size_t tcgen05_bt_bytes(int dtype, unsigned k, unsigned m, int flags, const Tuning &t) {
  if (tcgen05_b_in_place(dtype, flags, t)) return 0;
  const size_t eb = elem_bytes(dtype);
  return align_up(size_t(m) * k * eb * (split3(dtype, flags) ? 3 : 1), 1024);
}

size_t tcgen05_scratch_bytes(int dtype, unsigned n, unsigned k, unsigned m, int flags, const Tuning &t) {
  const size_t eb = elem_bytes(dtype);
  size_t bytes = TAIL_BYTES + tcgen05_bt_bytes(dtype, k, m, flags, t);  // counters (tail) + B copy
  if (dtype == MM_DTYPE_FLOAT || (flags & MM_FLAG_TRANSPOSED_A)) {
    bytes += align_up(size_t(n) * k * eb * (split3(dtype, flags) ? 3 : 1), 1024);
  }
  return bytes;
}

// Copy row-sliced B (slices on peer GPUs) into one local array: the NVLink all-gather of the
// multi-GPU path for the kernel families that consume B as stored.
int gather_b_rows(const BSource &src, void *dst, size_t elem_bytes, unsigned k, unsigned m, cudaStream_t stream) {
  return launch_panels(false, src, dst, elem_bytes, k, m, /*panel_cols=*/unsigned(1024 / elem_bytes), nullptr,
                       num_sms() * 2, stream);
}

int tcgen05_prepare_b(int dtype, const BSource &src, void *bt, unsigned k, unsigned m, int flags, const Tuning &t,
                      const void **b_op, unsigned int *ready, unsigned *ready_target, cudaStream_t stream) {
  *b_op = bt;
  if (ready_target) *ready_target = 0;
  const bool parts = src.src != nullptr;
  const size_t eb = elem_bytes(dtype);
  if (tcgen05_b_mn(dtype, flags, t)) {
    const bool in_place = tcgen05_b_in_place(dtype, flags, t);
    if (in_place && !parts) {
      *b_op = src.b;  // nothing to prepare
      return MM_OK;
    }
    // float: rounded copy (same layout).  half / unrounded float with slices: plain gather into `bt`.
    if (!in_place && !parts && ready == nullptr) {
      // one local array, nobody waiting on panels: the flat elementwise pass (6.3 TB/s against the panel
      // kernel's 5.1 on a 512 MiB block — the panel order costs row-segment locality)
      MM_CUDA_TRY(cudaFuncSetAttribute(round_tf32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
      const size_t count4 = size_t(k) * m / 4;
      const int blocks = int(std::min<size_t>((count4 + 255) / 256, size_t(num_sms()) * 16));
      round_tf32_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float4 *>(src.b), static_cast<float4 *>(bt), count4);
      MM_CUDA_TRY(cudaGetLastError());
      return MM_OK;
    }
    const unsigned panel_cols = unsigned(t.block_n());
    const unsigned panels = ceil_div(m, panel_cols);
    const bool publish = ready != nullptr && panels <= B_READY_BYTES / sizeof(unsigned int);
    if (publish && ready_target) *ready_target = ceil_div(k, PANEL_ROWS);
    // co-resident persistent grid (one 512-thread CTA per SM next to the GEMM's CTA) when the GEMM
    // consumes panels while this runs; a wider grid when it runs alone in stream order
    const int grid = publish ? num_sms() : num_sms() * 2;
    // An SM changes its L1 / shared-memory split only when it is idle.  This kernel uses no shared memory; were it
    // to run under the default (L1-heavy) split, the GEMM's CTAs (214 KiB of shared memory) could not become
    // resident next to it and would wait for it to END — measured: the "overlapped" GEMM took exactly its own time
    // plus this kernel's.  Ask for the shared-memory-heavy split so that both fit on an SM together.
    // (Function attributes are per device: set on every call, it is cheap.  A's rounding kernel may share SMs with
    // this one, so it asks for the same split — tcgen05_prepare_a.)
    MM_CUDA_TRY(cudaFuncSetAttribute(prep_b_panels_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    MM_CUDA_TRY(cudaFuncSetAttribute(prep_b_panels_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    return launch_panels(!in_place, src, bt, eb, k, m, panel_cols, publish ? ready : nullptr, grid, stream);
  }
  // K-major copy B^T (M x K): tuning knob b_mn = 0, and always for the 3xTF32 split
  const void *b = src.b;
  if (parts) return fail(MM_ERR_UNSUPPORTED, "row-sliced B needs the MN-major B path (gather it first)");
  if (split3(dtype, flags)) {
    dim3 grid((m + 63) / 64, (k + 63) / 64);
    split3_transpose_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const float *>(b), static_cast<float *>(bt), k, m);
  } else if (dtype == MM_DTYPE_FLOAT) {
    if (t.tf32_no_round()) {
      launch_transpose<float, false>(b, bt, k, m, stream);
    } else {
      launch_transpose<float, true>(b, bt, k, m, stream);
    }
  } else if (dtype == MM_DTYPE_UINT8) {
    launch_transpose<unsigned char, false>(b, bt, k, m, stream);
  } else {
    launch_transpose<__half, false>(b, bt, k, m, stream);
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

// `rows` rows of A -> the K-major A operand.  Row-major float A is rounded into `aprep`; row-major
// half A is used in place; A stored K x N (`transposed`, leading dimension = rows, only whole
// matrices) is transposed into `aprep`.  *a_op receives the operand pointer.
int tcgen05_prepare_a(int dtype, const void *a, void *aprep, unsigned rows, unsigned k, int flags, const Tuning &t,
                      const void **a_op, cudaStream_t stream) {
  const bool transposed = (flags & MM_FLAG_TRANSPOSED_A) != 0;
  *a_op = a;
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
      if (t.tf32_no_round()) {
        launch_transpose<float, false>(a, aprep, k, rows, stream);
      } else {
        launch_transpose<float, true>(a, aprep, k, rows, stream);  // A stored K x N -> N x K
      }
      *a_op = aprep;
    } else if (!t.tf32_no_round()) {
      MM_CUDA_TRY(cudaFuncSetAttribute(round_tf32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));  // see tcgen05_prepare_b
      const size_t count4 = size_t(rows) * k / 4;
      const int blocks = int(std::min<size_t>((count4 + 255) / 256, size_t(num_sms()) * 16));
      round_tf32_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float4 *>(a),
                                                   static_cast<float4 *>(aprep), count4);
      *a_op = aprep;
    }
  } else if (transposed) {
    if (dtype == MM_DTYPE_UINT8) launch_transpose<unsigned char, false>(a, aprep, k, rows, stream);
    else launch_transpose<__half, false>(a, aprep, k, rows, stream);
    *a_op = aprep;
  }
  MM_CUDA_TRY(cudaGetLastError());
  return MM_OK;
}

namespace {

struct LaunchPlan {
  const CUtensorMap *map_a, *map_b, *map_c;
  void *c;
  GemmParams p;
  int requested_stages;
  bool attributes_only;  // dry run: set the function attribute (loads the kernel), launch nothing
  cudaStream_t stream;
};

template <int KIND, typename TOut, int CG, int BN, bool BMN>
int launch_gemm_variant(LaunchPlan plan) {
  using G = Geo<CG, BN>;
  auto kern = gemm_tcgen05_kernel<KIND, TOut, CG, BN, BMN>;
  // Ring depth: the deepest that fits unless the tuning asks for less.  Measured (float 16384^3 /
  // half 32768^3, CTA pairs, WITH the soft wave barrier): depth 4 leaves the tensor pipe 79-86 %
  // active, depth 5-6 reach 96-98 % at unchanged DRAM traffic.
  const int stages = plan.requested_stages <= 0 ? G::MAX_STAGES
                                                 : std::min(std::max(plan.requested_stages, 2), int(G::MAX_STAGES));
  const size_t smem = G::smem_bytes(stages);
  MM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G::smem_bytes(G::MAX_STAGES))));
  if (plan.attributes_only) return MM_OK;
  plan.p.num_stages = uint32_t(stages);
  plan.p.raster_group = std::max<uint32_t>(1u, plan.p.raster_group / G::TILE_ROWS);  // rows -> row tiles
  const uint32_t tiles = ceil_div(plan.p.rows, G::TILE_ROWS) * ceil_div(plan.p.cols, BN);
  const uint32_t groups = std::min<uint32_t>(tiles, uint32_t(num_sms()) / CG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = plan.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (plan.p.tile_sync) MM_CUDA_TRY(cudaMemsetAsync(plan.p.tile_sync, 0, sizeof(unsigned int), plan.stream));
  MM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, *plan.map_a, *plan.map_b, *plan.map_c, static_cast<TOut *>(plan.c), plan.p));
  return MM_OK;
}

template <int KIND, typename TOut>
int dispatch_variant(int cg, int bn, bool bmn, const LaunchPlan &plan) {
#define MM_VARIANT(CGV, BNV, BMNV) \
  if (cg == CGV && bn == BNV && bmn == BMNV) return launch_gemm_variant<KIND, TOut, CGV, BNV, BMNV>(plan);
  MM_VARIANT(2, 256, true)
  MM_VARIANT(2, 256, false)
  MM_VARIANT(1, 256, true)
  MM_VARIANT(1, 256, false)
  MM_VARIANT(2, 128, true)
  MM_VARIANT(2, 128, false)
  MM_VARIANT(1, 128, true)
  MM_VARIANT(1, 128, false)
#undef MM_VARIANT
  return fail(MM_ERR_INVALID, "no tcgen05 kernel variant for this tuning (cta_group 1|2, block_n 128|256)");
}

int gemm_dispatch(int dtype, const void *a_op, const void *b_op, void *c, unsigned rows, unsigned k, unsigned m,
                  int flags, const Tuning &t, unsigned int *tile_sync, const unsigned int *b_ready,
                  unsigned b_ready_target, bool attributes_only, cudaStream_t stream) {
  if (split3(dtype, flags)) k *= 3;  // the operands carry [hi|hi|lo] x [hi|lo|hi] per 16-block of K
  const bool is_f32 = dtype == MM_DTYPE_FLOAT;
  const size_t eb = elem_bytes(dtype);
  const int cg = t.cta_group(), bn = t.block_n();
  const bool bmn = tcgen05_b_mn(dtype, flags, t);
  CUtensorMap map_a, map_b, map_c;
  std::memset(&map_c, 0, sizeof(map_c));
  LaunchPlan plan{&map_a, &map_b, &map_c, c, {}, t.stages(), attributes_only, stream};
  if (!attributes_only) {
    int rc = make_operand_map(&map_a, a_op, dtype, rows, k, BLOCK_M);
    if (rc != MM_OK) return rc;
    rc = bmn ? make_b_mn_map(&map_b, b_op, dtype, k, m) : make_operand_map(&map_b, b_op, dtype, m, k, uint32_t(bn / cg));
    if (rc != MM_OK) return rc;
    if (t.tma_store()) {
      rc = make_c_map(&map_c, c, dtype, rows, m);
      if (rc != MM_OK) return rc;
    }
  }
  plan.p.rows = rows;
  plan.p.cols = m;
  plan.p.k_bytes = uint32_t(size_t(k) * eb);
  plan.p.raster_group = uint32_t(std::max(1, t.raster_rows()));  // in rows here; per-variant tiles in the launcher
  plan.p.tma_store = t.tma_store() ? 1u : 0u;
  plan.p.b_ready_target = b_ready_target;
  plan.p.l2_policy = t.l2_policy() == 1 ? ptx::L2_EVICT_FIRST : (t.l2_policy() == 2 ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL);
  plan.p.tile_sync = t.tile_sync() ? tile_sync : nullptr;
  plan.p.b_ready = b_ready;
  if (dtype == MM_DTYPE_UINT8) {
    // MN-major 8-bit atoms are 128 columns wide: a CTA must stage at least one (tile columns / CTAs per group >= 128)
    if (bmn && bn / cg < 128) return dispatch_variant<ptx::KIND_I8, unsigned char>(cg, 256, bmn, plan);
    return dispatch_variant<ptx::KIND_I8, unsigned char>(cg, bn, bmn, plan);
  }
  return is_f32 ? dispatch_variant<ptx::KIND_TF32, float>(cg, bn, bmn, plan)
                : dispatch_variant<ptx::KIND_F16, __half>(cg, bn, bmn, plan);
}

}  // namespace

// C[rows x m] = Aop[rows x k] * B on the tensor cores; `b_op` as returned by tcgen05_prepare_b.
int tcgen05_gemm(int dtype, const void *a_op, const void *b_op, void *c, unsigned rows, unsigned k, unsigned m,
                 int flags, const Tuning &t, unsigned int *tile_sync, const unsigned int *b_ready,
                 unsigned b_ready_target, cudaStream_t stream) {
  return gemm_dispatch(dtype, a_op, b_op, c, rows, k, m, flags, t, tile_sync, b_ready, b_ready_target, false, stream);
}

int tcgen05_prepare_b_async(int dtype, const BSource &src, void *local_b, void *scratch, size_t scratch_bytes,
                            unsigned k, unsigned m, int flags, const Tuning &t, cudaStream_t stream, cudaStream_t side,
                            cudaEvent_t ev_fork, cudaEvent_t ev_join, PreparedB *out) {
  *out = PreparedB{};
  const bool in_place = tcgen05_b_in_place(dtype, flags, t);
  const bool parts = src.src != nullptr;
  if (parts && !tcgen05_b_mn(dtype, flags, t)) {
    // K-major copy requested (tuning / 3xTF32): assemble the slices first, then transpose locally
    int rc = gather_b_rows(src, local_b, elem_bytes(dtype), k, m, stream);
    if (rc != MM_OK) return rc;
    BSource whole;
    whole.b = local_b;
    return tcgen05_prepare_b(dtype, whole, scratch, k, m, flags, t, &out->b_op, nullptr, nullptr, stream);
  }
  void *bt = in_place ? local_b : scratch;
  const Tcgen05Counters cnt = tcgen05_counters(scratch, scratch_bytes);
  const unsigned panels = ceil_div(m, unsigned(t.block_n()));
  // the panel kernel runs (float rounding, or a gather of slices), a second stream exists, the tuning allows it
  const bool overlap = side != nullptr && t.b_overlap() != 0 && tcgen05_b_mn(dtype, flags, t) && (!in_place || parts) &&
                       panels <= B_READY_BYTES / sizeof(unsigned int);
  if (!overlap) return tcgen05_prepare_b(dtype, src, bt, k, m, flags, t, &out->b_op, nullptr, nullptr, stream);
  MM_CUDA_TRY(cudaMemsetAsync(cnt.b_ready, 0, panels * sizeof(unsigned int), stream));
  MM_CUDA_TRY(cudaEventRecord(ev_fork, stream));
  MM_CUDA_TRY(cudaStreamWaitEvent(side, ev_fork, 0));
  const int rc = tcgen05_prepare_b(dtype, src, bt, k, m, flags, t, &out->b_op, cnt.b_ready, &out->ready_target, side);
  cudaEventRecord(ev_join, side);  // the side stream rejoins whatever happened above
  out->forked = true;
  out->ready = cnt.b_ready;
  return rc;
}

int launch_tcgen05(int dtype, const GemmArgs &g, void *scratch, size_t scratch_bytes) {
  if (dtype != MM_DTYPE_FLOAT && dtype != MM_DTYPE_HALF && dtype != MM_DTYPE_UINT8) {
    return fail(MM_ERR_UNSUPPORTED, "tcgen05 path handles float, half and uint8_t only");
  }
  if (g.tuning == nullptr) return fail(MM_ERR_INVALID, "tcgen05 launch without tuning");
  const Tuning &t = *g.tuning;
  if (g.dry_run) {
    // force the lazily loaded kernels in (prep + the GEMM variant this tuning selects) and the driver entry point
    cudaFuncAttributes attr;
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, round_tf32_kernel));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, prep_b_panels_kernel<true>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, prep_b_panels_kernel<false>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, transpose_prep_kernel<float, true>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, transpose_prep_kernel<__half, false>));
    MM_CUDA_TRY(cudaFuncGetAttributes(&attr, transpose_prep_kernel<unsigned char, false>));
    if (!get_encode_fn()) return fail(MM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return gemm_dispatch(dtype, nullptr, nullptr, nullptr, g.n, g.k, g.m, g.flags, t, nullptr, nullptr, 0, true, g.stream);
  }
  if (scratch_bytes < tcgen05_scratch_bytes(dtype, g.n, g.k, g.m, g.flags, t)) {
    return fail(MM_ERR_INVALID, "tcgen05 scratch too small");
  }
  unsigned char *sp = static_cast<unsigned char *>(scratch);
  void *aprep = sp + tcgen05_bt_bytes(dtype, g.k, g.m, g.flags, t);
  const Tcgen05Counters cnt = tcgen05_counters(scratch, scratch_bytes);
  BSource src;
  src.b = g.b;
  PreparedB pb;
  const void *a_op = nullptr;
  int rc = tcgen05_prepare_b_async(dtype, src, nullptr, scratch, scratch_bytes, g.k, g.m, g.flags, t, g.stream,
                                   g.side_stream, g.ev_fork, g.ev_join, &pb);
  if (rc == MM_OK) rc = tcgen05_prepare_a(dtype, g.a, aprep, g.n, g.k, g.flags, t, &a_op, g.stream);
  if (rc == MM_OK && g.ev_prep_done) cudaEventRecord(g.ev_prep_done, g.stream);
  if (rc == MM_OK) {
    rc = tcgen05_gemm(dtype, a_op, pb.b_op, g.c, g.n, g.k, g.m, g.flags, t, cnt.tile_sync, pb.ready, pb.ready_target,
                      g.stream);
  }
  if (pb.forked) cudaStreamWaitEvent(g.stream, g.ev_join, 0);  // join, on the error paths too
  return rc;
}

}  // namespace mm
