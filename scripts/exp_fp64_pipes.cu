// Microbenchmark (registers only): throughput of the FP64 tensor instruction (DMMA.8x8x4) and of the
// FP64 CUDA-core FMA (DFMA) on one B200, alone and mixed, to find what bounds gemm_dmma_kernel:
//   * DMMA alone with 1 / 2 / 4 warps per scheduler  -> is 88 % pipe activity an issue-interval limit?
//   * DFMA alone                                      -> the vector FP64 rate
//   * DMMA warps + DFMA warps on the same SM          -> do the two share one datapath?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_fp64_pipes scripts/exp_fp64_pipes.cu
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// warps [0, dmma_warps) run DMMA chains, warps [dmma_warps, blockDim/32) run DFMA chains
template <int ACC>
__global__ void __launch_bounds__(1024, 1) mix_kernel(double *out, int iters, int dmma_warps, double seed) {
  const int warp = threadIdx.x / 32;
  double acc[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i][0] = acc[i][1] = seed * (i + 1);
  double a0 = seed + threadIdx.x * 1e-9, a1 = a0 * 0.5, b0 = 1.0 - a0, b1 = b0 * 0.25;
  if (warp < dmma_warps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ACC; ++i) dmma(acc[i][0], acc[i][1], (i & 1) ? a1 : a0, (i & 2) ? b1 : b0);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ACC; ++i) {
        acc[i][0] = fma(acc[i][0], a0, b0);
        acc[i][1] = fma(acc[i][1], a1, b1);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[threadIdx.x] = s;  // keep the chains alive
}

template <int ACC>
static void run(const char *label, int warps, int dmma_warps, int ctas_per_sm_hint) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *out;
  cudaMalloc(&out, 1024 * sizeof(double));
  const int iters = 1 << 15;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * ctas_per_sm_hint;
  mix_kernel<ACC><<<grid, warps * 32, 0>>>(out, 64, dmma_warps, 0.001);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mix_kernel<ACC><<<grid, warps * 32, 0>>>(out, iters, dmma_warps, 0.001);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dmma_fma = double(grid) * dmma_warps * double(iters) * ACC * 256.0;
  const double dfma_fma = double(grid) * (warps - dmma_warps) * double(iters) * ACC * 2.0 * 32.0;
  printf("{\"case\": \"%s\", \"acc\": %d, \"warps\": %d, \"dmma_warps\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, "
         "\"dfma_tflops\": %.2f, \"total_tflops\": %.2f, \"err\": \"%s\"}\n",
         label, ACC, warps, dmma_warps, ms, 2e-9 * dmma_fma / ms, 2e-9 * dfma_fma / ms,
         2e-9 * (dmma_fma + dfma_fma) / ms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<16>("dmma_only", 4, 4, 1);
  run<16>("dmma_only", 8, 8, 1);
  run<16>("dmma_only", 16, 16, 1);
  run<32>("dmma_only", 8, 8, 1);
  run<8>("dmma_only", 8, 8, 1);
  run<16>("dfma_only", 8, 0, 1);
  run<16>("dfma_only", 16, 0, 1);
  run<16>("dfma_only", 32, 0, 1);
  run<16>("mixed", 16, 8, 1);
  run<16>("mixed", 12, 8, 1);
  run<16>("mixed", 24, 8, 1);
  run<16>("mixed", 32, 16, 1);
  return 0;
}
