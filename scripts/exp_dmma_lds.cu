// Microbenchmark: the LDS + DMMA inner loop of gemm_dmma_kernel on a resident shared-memory tile
// (no global traffic, no barriers), 8 warps of 64 x 32 per CTA, one CTA per SM.
//   variant 0: no LDS (operands stay in registers)            -> DMMA pipe ceiling
//   variant 1: LDS.64 fragments, pitches 36 / 132 (the kernel as shipped in r01)
//   variant 2: LDS.128 fragments: k-step pairs for A, column pairs for B, pitches 40 / 130
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_dmma_lds scripts/exp_dmma_lds.cu
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int BM = 128, BN = 128, BK = 32, MI = 8, NJ = 4, WN = 4;

template <int VARIANT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) loop_kernel(double *out, int k_tiles) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LDA = VARIANT == 2 ? 40 : 36, LDB = VARIANT == 2 ? 130 : 132;
  double *As = smem, *Bs = smem + BM * LDA;
  for (int i = threadIdx.x; i < BM * LDA + BK * LDB; i += blockDim.x) smem[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int wr = (warp / WN) % 2, wc = warp % WN, g = lane / 4, q = lane % 4;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < k_tiles; ++kt) {
    // defeat hoisting: the tile base depends on kt through an opaque zero
    int zero;
    asm volatile("mov.u32 %0, 0;" : "=r"(zero) : "r"(kt));
    const double *as = As + zero, *bs = Bs + zero;
    if (VARIANT == 0) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = as[(wr * 64 + i * 8 + g) * LDA + q];
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = bs[q * LDB + wc * 32 + j * 8 + g];
#pragma unroll
      for (int k4 = 0; k4 < BK; k4 += 4)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    } else if (VARIANT == 1) {
#pragma unroll
      for (int k4 = 0; k4 < BK; k4 += 4) {
        double af[MI], bf[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[i] = as[(wr * 64 + i * 8 + g) * LDA + k4 + q];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[j] = bs[(k4 + q) * LDB + wc * 32 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
#pragma unroll
      for (int s2 = 0; s2 < BK / 8; ++s2) {
        double2 af[MI];       // .x -> step 2*s2, .y -> step 2*s2+1 (physical k = s2*8 + 2q + e)
        double2 bf[2][NJ / 2];  // [e][h]: tiles j = 2h, 2h+1 (physical cols wc*32 + 16h + 2g + {0,1})
#pragma unroll
        for (int i = 0; i < MI; ++i)
          af[i] = *reinterpret_cast<const double2 *>(as + (wr * 64 + i * 8 + g) * LDA + s2 * 8 + 2 * q);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int h = 0; h < NJ / 2; ++h)
            bf[e][h] = *reinterpret_cast<const double2 *>(bs + (s2 * 8 + 2 * q + e) * LDB + wc * 32 + 16 * h + 2 * g);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int h = 0; h < NJ / 2; ++h) {
              dmma(acc[i][2 * h][0], acc[i][2 * h][1], e ? af[i].y : af[i].x, bf[e][h].x);
              dmma(acc[i][2 * h + 1][0], acc[i][2 * h + 1][1], e ? af[i].y : af[i].x, bf[e][h].y);
            }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VARIANT, int WARPS>
static void run(const char *label) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, size_t(sms) * WARPS * 32 * sizeof(double));
  const size_t smem = (BM * 40 + BK * 132) * sizeof(double);
  cudaFuncSetAttribute(loop_kernel<VARIANT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int k_tiles = 4096;
  loop_kernel<VARIANT, WARPS><<<sms, WARPS * 32, smem>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  loop_kernel<VARIANT, WARPS><<<sms, WARPS * 32, smem>>>(out, k_tiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = double(sms) * WARPS * double(k_tiles) * (BK / 4) * MI * NJ * 256.0;
  printf("{\"case\": \"%s\", \"variant\": %d, \"warps\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", label,
         VARIANT, WARPS, ms, 2e-9 * fma / ms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0, 8>("registers_only");
  run<1, 8>("lds64_pitch36_132");
  run<2, 8>("lds128_pitch40_130");
  run<1, 8>("lds64_pitch36_132");
  run<2, 8>("lds128_pitch40_130");
  return 0;
}
