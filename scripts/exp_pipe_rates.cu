// Issue / pipe rates behind the float (Add, Min) semiring bound, registers only — no shared
// memory, no barriers, 256 threads x 2 CTAs per SM like semiring_tile_kernel (4 warps per scheduler):
//   fadd      scalar add.rn.f32                         (FMA pipe)
//   fadd2     packed add.rn.f32x2 = FADD2               (FMA pipe, two results per lane)
//   fmnmx     2-input min.f32 = FMNMX                   (ALU pipe)
//   fmnmx3    3-input min.f32 = FMNMX3                  (ALU pipe)
//   mix       the kernel's inner step: 2 FADD2 feed 2 FMNMX3 (4 element-steps per 4 instructions)
//   mix_swap  the same with the second pair swapped (sources of an FMNMX3 in registers of opposite parity)
// Every operation is an `asm volatile`, so ptxas neither merges nor reorders across chains; 32 independent
// chains per thread hide the 4-cycle dependent-issue latency.  Output: cycles per warp instruction per
// scheduler at the measured clock (clock64 of one warp), and the TOp/s the (Add, Min) kernel would reach if
// that were its only limit (2 ops per element-step, 148 SMs x 4 schedulers x 32 lanes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_pipe_rates scripts/exp_pipe_rates.cu
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

enum { FADD, FADD2, FMNMX, FMNMX3, MIX, MIX_SWAP, MIX_LDS };

template <int MODE>
__global__ void __launch_bounds__(256, 2) rate_kernel(float *out, int iters, long long *cycles) {
  __shared__ __align__(16) float tile[32 * 128];
  if (MODE == MIX_LDS) {
    for (int i = threadIdx.x; i < 32 * 128; i += 256) tile[i] = 1e-3f * (i % 97);
    __syncthreads();
  }
  float acc[64], a[8], b[8];
  unsigned long long ap[8], bp[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 3.0e38f - i;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = 1e-3f * (threadIdx.x + i);
    b[i] = 2e-3f * (threadIdx.x * 3 + i);
    ap[i] = pack2(a[i], a[i]);
    bp[i] = pack2(b[i], b[(i + 1) % 8]);
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float &x = acc[i * 8 + 2 * p], &y = acc[i * 8 + 2 * p + 1];
        if (MODE == FADD) {
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(b[p]));
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(y) : "f"(b[p + 4]));
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(a[i]));
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(y) : "f"(a[i]));
        } else if (MODE == FADD2) {
          unsigned long long v = pack2(x, y);
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bp[p]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(ap[i]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bp[p + 4]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(ap[(i + 1) % 8]));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
        } else if (MODE == FMNMX) {
          asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(b[p]));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(y) : "f"(b[p + 4]));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(a[i]));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(y) : "f"(a[i]));
        } else if (MODE == FMNMX3) {
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(b[p]), "f"(a[i]));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(b[p + 4]), "f"(a[i]));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a[(i + 1) % 8]), "f"(b[p]));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(a[(i + 1) % 8]), "f"(b[p + 4]));
        } else {
          // 2 FADD2 (k and k+1, columns 2p / 2p+1) -> 2 FMNMX3: four element-steps in four instructions.
          // The A pair is loop-carried (rotated through the chain below), so no add is loop-invariant.
          unsigned long long s0, s1;
          float s0l, s0h, s1l, s1h;
          asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(s0) : "l"(ap[i]), "l"(bp[p]));
          asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(s1) : "l"(ap[(i + 1) % 8]), "l"(bp[p + 4]));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(s0l), "=f"(s0h) : "l"(s0));
          if (MODE == MIX_SWAP) {
            asm("mov.b64 {%0, %1}, %2;" : "=f"(s1h), "=f"(s1l) : "l"(s1));
          } else {
            asm("mov.b64 {%0, %1}, %2;" : "=f"(s1l), "=f"(s1h) : "l"(s1));
          }
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(s0l), "f"(s1l));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(s0h), "f"(s1h));
        }
      }
      if (MODE >= MIX) {   // one extra FADD2 per 16 timed instructions: keeps every add of the next iteration new
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(ap[i]) : "l"(bp[i]));
      }
    }
    if (MODE == MIX_LDS) {   // the kernel's fragment traffic: 8 LDS.128 per 128 math instructions
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4 *>(tile + ((it * 8 + q) % 32) * 128 + (threadIdx.x % 16) * 4 + (q & 1) * 64);
        bp[q] = pack2(v.x + v.z, v.y + v.w);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char *label, double element_steps_per_instr) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 2, iters = 20000;
  float *out;
  long long *cyc, h_cyc = 0;
  cudaMalloc(&out, size_t(grid) * 256 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate_kernel<MODE><<<grid, 256>>>(out, 100, cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  rate_kernel<MODE><<<grid, 256>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h_cyc, cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost);
  // 128 timed instructions per thread and iteration; in the mix modes + the 8 chain-advancing packed adds, which
  // ptxas emits as 16 scalar FADD (SASS: 64 FADD2 + 64 FMNMX3 + 16 FADD per iteration) = 144 issue slots; the LDS
  // mode's 8 LDS.128 and the 16 adds that consume them are overhead on top, not counted.  16 warps per SM = 4 per scheduler
  const double instr_per_sched = 4.0 * (MODE >= MIX ? 144.0 : 128.0) * iters;
  const double cyc_per_instr = double(h_cyc) / instr_per_sched;
  const double mhz = double(h_cyc) / (ms * 1e3);
  const double tops_at_1965 = 2.0 * element_steps_per_instr * 32 * 4 * sms * 1965e6 / cyc_per_instr * 1e-12;
  printf("{\"case\": \"%s\", \"ms\": %.3f, \"sm_mhz\": %.0f, \"cycles_per_warp_instr_per_scheduler\": %.3f, "
         "\"addmin_tops_if_only_limit_at_1965MHz\": %.1f, \"err\": \"%s\"}\n",
         label, ms, mhz, cyc_per_instr, tops_at_1965, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  // element-steps one instruction accounts for in the (Add, Min) kernel: an FADD2 covers 2 of the adds, an FMNMX3 2 of
  // the mins; in the mix 4 instructions cover 4 element-steps
  run<FADD>("fadd", 1.0);
  run<FADD2>("fadd2", 2.0);
  run<FMNMX>("fmnmx", 1.0);
  run<FMNMX3>("fmnmx3", 2.0);
  run<MIX>("mix_2fadd2_2fmnmx3", 1.0);
  run<MIX_SWAP>("mix_swapped_second_pair", 1.0);
  run<MIX_LDS>("mix_plus_8_lds128_per_128", 1.0);
  return 0;
}
