// Microbenchmark for the (Add, Min) float inner loop of semiring_tile_kernel: per pair of k-steps and
// per accumulator 2 FADD + 1 three-input FMNMX (1.5 issue slots per element-step; ceiling 49.6 TOp/s at
// 1965 MHz).  256 threads, 8 x 8 accumulators per thread, 2 CTAs per SM as in the kernel.
//   variant 0: operands in registers; ptxas merges the repeated adds, leaving 512 FMNMX3 + 192 FADD per
//              k-tile: this measures the FMNMX3 (ALU pipe) rate, reported as if all 1536 slots had issued
//   variant 1: fragments read with LDS.128 from a resident tile, no barriers     -> + fragment traffic
//   variant 2: variant 1 + a __syncthreads() every 16 k-steps (the kernel's k-tile boundary)
//   variant 3: packed adds (add.rn.f32x2 -> FADD2, two adjacent columns per instruction): 1 FADD2 + 1 FMNMX3
//              per two element-steps (1.0 issue slot per element-step; ceiling 74.4 TOp/s)
//   variant 4: variant 3 with the column pair as the outer loop (sensitivity to instruction order)
//   variant 5: variant 3 with the B pair of the second k swapped (free: FADD2 takes an .F32x2.LO_HI operand),
//              so that the two sums an FMNMX3 reduces sit in registers of opposite parity (in variant 3 all
//              three FMNMX3 sources have the same parity: candidates for register-bank conflicts)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_semiring_issue scripts/exp_semiring_issue.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int BK = 16, LDA = 132, LDB = 128;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void add2(unsigned long long a, unsigned long long b, float &lo, float &hi) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 2) loop_kernel(float *out, int k_tiles) {
  __shared__ __align__(16) float As[BK * LDA];
  __shared__ __align__(16) float Bs[BK * LDB];
  for (int i = threadIdx.x; i < BK * LDA; i += 256) As[i] = 1e-3f * (i % 89);
  for (int i = threadIdx.x; i < BK * LDB; i += 256) Bs[i] = 2e-3f * (i % 83);
  __syncthreads();
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 3.0e38f;

  float af[2][8], bf[2][8];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      af[u][q] = As[u * LDA + ty * 4 + (q / 4) * 64 + q % 4];
      bf[u][q] = Bs[u * LDB + tx * 4 + (q / 4) * 64 + q % 4];
    }

  for (int kt = 0; kt < k_tiles; ++kt) {
    int zero;
    asm volatile("mov.u32 %0, 0;" : "=r"(zero) : "r"(kt));
    const float *as = As + zero, *bs = Bs + zero;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 2) {
      if (VARIANT >= 1) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 a0 = *reinterpret_cast<const float4 *>(as + (kk + u) * LDA + ty * 4);
          const float4 a1 = *reinterpret_cast<const float4 *>(as + (kk + u) * LDA + 64 + ty * 4);
          const float4 b0 = *reinterpret_cast<const float4 *>(bs + (kk + u) * LDB + tx * 4);
          const float4 b1 = *reinterpret_cast<const float4 *>(bs + (kk + u) * LDB + 64 + tx * 4);
          af[u][0] = a0.x, af[u][1] = a0.y, af[u][2] = a0.z, af[u][3] = a0.w;
          af[u][4] = a1.x, af[u][5] = a1.y, af[u][6] = a1.z, af[u][7] = a1.w;
          bf[u][0] = b0.x, bf[u][1] = b0.y, bf[u][2] = b0.z, bf[u][3] = b0.w;
          bf[u][4] = b1.x, bf[u][5] = b1.y, bf[u][6] = b1.z, bf[u][7] = b1.w;
        }
      }
      if (VARIANT >= 3) {
        unsigned long long bp[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int p = 0; p < 4; ++p)
            bp[u][p] = (VARIANT == 5 && u == 1) ? pack2(bf[u][2 * p + 1], bf[u][2 * p]) : pack2(bf[u][2 * p], bf[u][2 * p + 1]);
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const int i = VARIANT != 4 ? x : (y * 2 + x / 4) % 8, p = VARIANT != 4 ? y : x % 4;
            float t0l, t0h, t1l, t1h;
            add2(pack2(af[0][i], af[0][i]), bp[0][p], t0l, t0h);
            if (VARIANT == 5) {
              add2(pack2(af[1][i], af[1][i]), bp[1][p], t1h, t1l);  // swapped pair: .x is column 2p+1
            } else {
              add2(pack2(af[1][i], af[1][i]), bp[1][p], t1l, t1h);
            }
            acc[i][2 * p] = fminf(fminf(acc[i][2 * p], t0l), t1l);
            acc[i][2 * p + 1] = fminf(fminf(acc[i][2 * p + 1], t0h), t1h);
          }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (VARIANT == 0) {
            float t0, t1;
            asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(t0) : "f"(af[0][i]), "f"(bf[0][j]));
            asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(t1) : "f"(af[1][i]), "f"(bf[1][j]));
            acc[i][j] = fminf(fminf(acc[i][j], t0), t1);
          } else {
            acc[i][j] = fminf(fminf(acc[i][j], __fadd_rn(af[0][i], bf[0][j])), __fadd_rn(af[1][i], bf[1][j]));
          }
        }
    }
    if (VARIANT == 2) __syncthreads();
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int VARIANT>
static void run(const char *label) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 2;
  float *out;
  cudaMalloc(&out, size_t(grid) * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int k_tiles = 8192;
  loop_kernel<VARIANT><<<grid, 256>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  loop_kernel<VARIANT><<<grid, 256>>>(out, k_tiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 2.0 * double(grid) * 256 * 64 * double(k_tiles) * BK;
  printf("{\"case\": \"%s\", \"variant\": %d, \"ms\": %.3f, \"tops\": %.2f, \"frac_of_49.6\": %.3f, \"err\": \"%s\"}\n", label, VARIANT,
         ms, 1e-9 * ops / ms, 1e-9 * ops / ms / 49.6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("registers_only");
  run<1>("lds128_fragments");
  run<2>("lds128_fragments_barrier_per_ktile");
  run<3>("packed_fadd2");
  run<4>("packed_fadd2_column_pair_outer");
  run<5>("packed_fadd2_second_k_swapped");
  run<1>("lds128_fragments");
  run<3>("packed_fadd2");
  run<5>("packed_fadd2_second_k_swapped");
  return 0;
}
