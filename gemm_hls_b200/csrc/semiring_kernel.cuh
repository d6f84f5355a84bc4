// CUDA-core semiring tile kernel:  C[N x M] = A[N x K] (x) B[K x M]  for ANY (Map, Reduce, T).
//
// B200 counterpart of the reference's ProcessingElement chain (kernel/Compute.cpp:11-231) fed by
// ReadA/TransposeA/ReadB/FeedB and drained by WriteC (kernel/Memory.cpp:58-438), for the
// configurations that are not a dense (Multiply, Add) contraction on a tensor-core type — and,
// under MM_FLAG_EXACT, for those too.  Like the reference it computes an outer product per k into
// an on-chip C tile that is written once; unlike the reference's literal-0 seed
// (kernel/Compute.cpp:116-118, which breaks Min — SURVEY.md section 5 trap 1) the accumulators
// start from OperatorReduce::identity(), the result definition of Naive<> (include/Utility.h:29).
//
// Exactness: every C element is reduced by ONE thread, sequentially over k = 0..K-1, with one
// rounding per Map and per Reduce (semiring.cuh), so the output is bit-identical to Naive<>.
//
// Tiling: CTA tile BM x BN = 128 x 128, k-step BK = 64 bytes of K (the reference's memory word,
// so K % BK == 0 is implied by the reference's own shape rule), 256 threads, 8 x 8 accumulators
// per thread laid out as 2 x 2 quads of 4 so that shared-memory fragment reads are 16-byte
// conflict-free and global C stores are row-contiguous.  A is transposed on the way into shared
// memory (the role of TransposeA, kernel/Memory.cpp:130-181); tiles are double buffered with the
// next tile's global loads in flight during the current tile's compute.
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "semiring.cuh"

namespace mm {

template <typename T>
struct alignas((sizeof(T) * 4 <= 16) ? sizeof(T) * 4 : 16) Quad {
  T v[4];
};

template <typename T>
struct alignas(16) Chunk16 {  // one 16-byte global load
  T v[16 / sizeof(T)];
};

template <typename T>
struct SemiringTile {
  static constexpr int BM = 128;
  static constexpr int BN = 128;
  static constexpr int BK = 64 / sizeof(T);   // elements of K per step (64 bytes)
  static constexpr int VEC = 16 / sizeof(T);  // elements per 16-byte chunk
  static constexpr int THREADS = 256;
  // 128 rows x 64 B of A and BK rows x 128 cols of B are both 512 16-byte chunks.
  static constexpr int CHUNKS = 512;
  static constexpr int CHUNKS_PER_THREAD = CHUNKS / THREADS;  // 2
  static constexpr int A_CHUNKS_PER_ROW = BK / VEC;           // 4
  static constexpr int B_CHUNKS_PER_ROW = BN / VEC;
  static constexpr int PAD = 4;                               // elements; keeps 16 B alignment for T >= 4 B
  static constexpr int LDA = BM + ((sizeof(T) >= 4) ? PAD : 16 / sizeof(T));
  static constexpr int LDB = BN;
  static constexpr size_t SMEM_BYTES = 2 * (size_t(BK) * LDA + size_t(BK) * LDB) * sizeof(T);
};

// 2 CTAs (16 warps) per SM for 4-byte element types: 64 accumulators + two k-steps of fragments fit
// in 128 registers without spilling.  8-byte types need the full 255-register budget, and 1- and
// 2-byte types (one 32-bit register per unpacked element) spill at 128: those run 1 CTA per SM.
template <typename T, class Map, class Reduce>
__global__ void __launch_bounds__(256, (sizeof(T) == 4) ? 2 : 1)
semiring_tile_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ C,
                     unsigned size_n, unsigned size_k, unsigned size_m,
                     bool TRANSPOSED_A) {
  using Cfg = SemiringTile<T>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, VEC = Cfg::VEC;
  constexpr int LDA = Cfg::LDA, LDB = Cfg::LDB;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *As = reinterpret_cast<T *>(smem_raw);                 // [2][BK][LDA]  (k-major: A transposed)
  T *Bs = As + 2 * BK * LDA;                               // [2][BK][LDB]

  const int tid = threadIdx.x;
  const int tx = tid % 16;  // column quad index
  const int ty = tid / 16;  // row quad index
  const size_t row0 = size_t(blockIdx.y) * BM;
  const size_t col0 = size_t(blockIdx.x) * BN;

  T acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = Reduce::identity();
  }

  Chunk16<T> a_stage[Cfg::CHUNKS_PER_THREAD];
  Chunk16<T> b_stage[Cfg::CHUNKS_PER_THREAD];

  auto load_global = [&](unsigned k0) {
#pragma unroll
    for (int i = 0; i < Cfg::CHUNKS_PER_THREAD; ++i) {
      const int c = tid + i * Cfg::THREADS;
      if (!TRANSPOSED_A) {
        // A row-major N x K: chunk = (row, 16-byte part of the 64-byte k-slab)
        const int r = c / Cfg::A_CHUNKS_PER_ROW;
        const int part = c % Cfg::A_CHUNKS_PER_ROW;
        size_t row = row0 + r;
        if (row >= size_n) row = size_n - 1;  // clamp: rows past N are computed but never stored
        a_stage[i] = *reinterpret_cast<const Chunk16<T> *>(A + row * size_k + k0 + part * VEC);
      } else {
        // A stored K x N: element-wise (N need not be a multiple of the vector width)
        const int kk = c / Cfg::B_CHUNKS_PER_ROW;
        const int part = c % Cfg::B_CHUNKS_PER_ROW;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          size_t row = row0 + part * VEC + v;
          if (row >= size_n) row = size_n - 1;
          a_stage[i].v[v] = A[size_t(k0 + kk) * size_n + row];
        }
      }
      {
        const int kk = c / Cfg::B_CHUNKS_PER_ROW;
        const int part = c % Cfg::B_CHUNKS_PER_ROW;
        size_t col = col0 + size_t(part) * VEC;
        if (col + VEC > size_m) col = size_m - VEC;  // M % VEC == 0 by the shape rule
        b_stage[i] = *reinterpret_cast<const Chunk16<T> *>(B + size_t(k0 + kk) * size_m + col);
      }
    }
  };

  auto store_shared = [&](int buf) {
    T *as = As + buf * BK * LDA;
    T *bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int i = 0; i < Cfg::CHUNKS_PER_THREAD; ++i) {
      const int c = tid + i * Cfg::THREADS;
      if (!TRANSPOSED_A) {
        const int r = c / Cfg::A_CHUNKS_PER_ROW;
        const int part = c % Cfg::A_CHUNKS_PER_ROW;
#pragma unroll
        for (int v = 0; v < VEC; ++v) as[(part * VEC + v) * LDA + r] = a_stage[i].v[v];
      } else {
        const int kk = c / Cfg::B_CHUNKS_PER_ROW;
        const int part = c % Cfg::B_CHUNKS_PER_ROW;
        *reinterpret_cast<Chunk16<T> *>(as + kk * LDA + part * VEC) = a_stage[i];
      }
      {
        const int kk = c / Cfg::B_CHUNKS_PER_ROW;
        const int part = c % Cfg::B_CHUNKS_PER_ROW;
        *reinterpret_cast<Chunk16<T> *>(bs + kk * LDB + part * VEC) = b_stage[i];
      }
    }
  };

  const unsigned k_tiles = size_k / BK;
  load_global(0);
  store_shared(0);
  __syncthreads();

  for (unsigned kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) load_global((kt + 1) * BK);

    const T *as = As + buf * BK * LDA;
    const T *bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 2) {
      // two consecutive k per step, reduced in order inside ONE expression
      //   acc = Reduce(Reduce(acc, Map(a_k, b_k)), Map(a_k+1, b_k+1))
      // so that a 3-input hardware reduction (FMNMX3 for float min/max) can be selected; the
      // evaluation order, and therefore every rounding, is the sequential order of Naive<>.
      T af[2][8], bf[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const Quad<T> a0 = *reinterpret_cast<const Quad<T> *>(as + (kk + u) * LDA + ty * 4);
        const Quad<T> a1 = *reinterpret_cast<const Quad<T> *>(as + (kk + u) * LDA + 64 + ty * 4);
        const Quad<T> b0 = *reinterpret_cast<const Quad<T> *>(bs + (kk + u) * LDB + tx * 4);
        const Quad<T> b1 = *reinterpret_cast<const Quad<T> *>(bs + (kk + u) * LDB + 64 + tx * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          af[u][q] = a0.v[q];
          af[u][4 + q] = a1.v[q];
          bf[u][q] = b0.v[q];
          bf[u][4 + q] = b1.v[q];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = Reduce::Apply(Reduce::Apply(acc[i][j], Map::Apply(af[0][i], bf[0][j])),
                                    Map::Apply(af[1][i], bf[1][j]));
        }
      }
    }

    if (kt + 1 < k_tiles) store_shared(buf ^ 1);
    __syncthreads();
  }

  // Write the C tile once, masked to n < N, m < M (the role of WriteC, kernel/Memory.cpp:361-392).
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t row = row0 + (i / 4) * 64 + ty * 4 + (i % 4);
    if (row >= size_n) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t col = col0 + h * 64 + tx * 4;
      if (col + 4 <= size_m) {
        Quad<T> out;
#pragma unroll
        for (int q = 0; q < 4; ++q) out.v[q] = acc[i][h * 4 + q];
        *reinterpret_cast<Quad<T> *>(C + row * size_m + col) = out;
      }
    }
  }
}

template <typename T, class Map, class Reduce>
int launch_semiring_typed(const void *a, const void *b, void *c, unsigned n, unsigned k, unsigned m,
                          bool transposed_a, cudaStream_t stream) {
  using Cfg = SemiringTile<T>;
  if (a == nullptr) {  // dry run: only make sure the kernel is loaded
    cudaFuncAttributes attr;
    return static_cast<int>(cudaFuncGetAttributes(&attr, semiring_tile_kernel<T, Map, Reduce>));
  }
  dim3 grid((m + Cfg::BN - 1) / Cfg::BN, (n + Cfg::BM - 1) / Cfg::BM);
  dim3 block(Cfg::THREADS);
  const T *pa = static_cast<const T *>(a);
  const T *pb = static_cast<const T *>(b);
  T *pc = static_cast<T *>(c);
  semiring_tile_kernel<T, Map, Reduce><<<grid, block, Cfg::SMEM_BYTES, stream>>>(pa, pb, pc, n, k, m,
                                                                                transposed_a);
  return static_cast<int>(cudaGetLastError());
}

// One translation unit per (data type, map operator) instantiates the five reduce operators
// (semiring_inst.cu compiled with -DMM_INST_T=<type> -DMM_INST_MAP=<MM_OP_*>); 30 small units
// build in parallel.
template <typename T, int MAP_OP>
int launch_semiring_for(int reduce_op, const void *a, const void *b, void *c, unsigned n, unsigned k,
                        unsigned m, bool ta, cudaStream_t stream);

#define MM_SEMIRING_CASE(REDOP)                                                                    \
  if (reduce_op == REDOP)                                                                          \
    return launch_semiring_typed<T, typename OpSelect<T, MAP_OP>::type,                            \
                                 typename OpSelect<T, REDOP>::type>(a, b, c, n, k, m, ta, stream);

#define MM_INSTANTIATE_SEMIRING(TYPE, MAPOP)                                                       \
  template <>                                                                                      \
  int launch_semiring_for<TYPE, MAPOP>(int reduce_op, const void *a, const void *b, void *c,       \
                                       unsigned n, unsigned k, unsigned m, bool ta,                \
                                       cudaStream_t stream) {                                      \
    using T = TYPE;                                                                                \
    constexpr int MAP_OP = MAPOP;                                                                  \
    MM_SEMIRING_CASE(MM_OP_MULTIPLY)                                                               \
    MM_SEMIRING_CASE(MM_OP_ADD)                                                                    \
    MM_SEMIRING_CASE(MM_OP_MIN)                                                                    \
    MM_SEMIRING_CASE(MM_OP_MAX)                                                                    \
    MM_SEMIRING_CASE(MM_OP_AND)                                                                    \
    if constexpr (std::is_same<T, float>::value) {                                                 \
      MM_SEMIRING_CASE(MM_OP_MIN_FAST)                                                             \
      MM_SEMIRING_CASE(MM_OP_MAX_FAST)                                                             \
    }                                                                                              \
    return -1;                                                                                     \
  }

}  // namespace mm
