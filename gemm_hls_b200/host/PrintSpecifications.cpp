// PrintSpecifications N K M [<SM clock MHz>] — counterpart of the reference's
// src/PrintSpecifications.cpp:16-80: pure arithmetic on the build configuration, here for the
// B200 kernels: operation count, the kernel family this configuration dispatches to, its peak
// model, the ideal runtime and the reference's own communication-volume model
// Q = N*M*(1 + K/T_N + K/T_M) elements (src/PrintSpecifications.cpp:72-78) with the CTA tile in
// the role of the FPGA memory tile.
#include <cstdlib>
#include <iostream>
#include <string>

#include "MatrixMultiplication.h"

static void PrintUsage(char **argv) {
#ifndef MM_DYNAMIC_SIZES
  std::cerr << "Usage: " << argv[0] << " [<SM clock MHz>]\n" << std::flush;
#else
  std::cerr << "Usage: " << argv[0] << " N K M [<SM clock MHz>]\n" << std::flush;
#endif
}

int main(int argc, char **argv) {
#ifdef MM_DYNAMIC_SIZES
  if (argc > 5 || argc < 4) {
    PrintUsage(argv);
    return 1;
  }
  const unsigned size_n = std::stoul(argv[1]);
  const unsigned size_k = std::stoul(argv[2]);
  const unsigned size_m = std::stoul(argv[3]);
  int next_arg = 4;
#else
  if (argc > 2) {
    PrintUsage(argv);
    return 1;
  }
  constexpr auto size_n = kSizeN;
  constexpr auto size_k = kSizeK;
  constexpr auto size_m = kSizeM;
  int next_arg = 1;
#endif
  float frequency = 1965.0f;  // B200 clocks.max.sm (MHz)
  if (argc > next_arg) frequency = std::stof(argv[next_arg]);

  constexpr int kSMs = 148;
  const std::string path = mm_kernel_path(kDataTypeCode, kMapOpCode, kReduceOpCode, kKernelFlags);
  // per-SM operations per cycle of the pipe each kernel family is bound by
  double ops_per_sm_clk;
  unsigned long tile_n, tile_m;
  if (path == "tcgen05_f16") {
    ops_per_sm_clk = 2.0 * 4096;  // 128x256x16 MACs per 128 cycles
    tile_n = 128; tile_m = 256;
  } else if (path == "tcgen05_tf32") {
    ops_per_sm_clk = 2.0 * 2048;  // 128x256x8 MACs per 128 cycles
    tile_n = 128; tile_m = 256;
  } else if (path == "dmma_f64") {
    ops_per_sm_clk = 2.0 * 64;    // FP64 DMMA ~ 64 FMA / clk / SM
    tile_n = 128; tile_m = 128;
  } else {
    ops_per_sm_clk = 2.0 * 64;    // map + reduce: two CUDA-core instructions per element-step, issue bound
    tile_n = 128; tile_m = 128;
  }
  const unsigned long long nOps = 2 * static_cast<unsigned long long>(size_n) * size_k * size_m;
  const double peak = 1e-3 * ops_per_sm_clk * kSMs * frequency;  // GOp/s
  const unsigned long tiles_n = (size_n + tile_n - 1) / tile_n, tiles_m = (size_m + tile_m - 1) / tile_m;
  const unsigned long tiles = tiles_n * tiles_m;
  const unsigned long waves = (tiles + kSMs - 1) / kSMs;
  const double ideal_runtime = 1e-9 * nOps / peak;
  // whole waves of full tiles: what the persistent schedule actually executes
  const double expected_runtime =
      static_cast<double>(waves) * (2.0 * tile_n * tile_m * size_k) / (ops_per_sm_clk * 1e6 * frequency);
  std::cout << "Configuration:        " << kDataTypeName << " (" << kMapOpName << ", " << kReduceOpName << ")\n";
  std::cout << "Kernel family:        " << path << "\n";
  std::cout << "Frequency:            " << frequency << " MHz\n";
  std::cout << "Number of operations: " << nOps << " (" << static_cast<float>(nOps) << ")\n";
  std::cout << "Expected runtime:     " << expected_runtime << " seconds\n";
  std::cout << "Ideal runtime:        " << ideal_runtime << " seconds\n";
  std::cout << "Percentage of ideal:  " << 100 * ideal_runtime / expected_runtime << "%\n";
  std::cout << "Expected performance: " << 1e-9 * nOps / expected_runtime << " GOp/s\n";
  std::cout << "Ideal performance:    " << peak << " GOp/s\n";
  std::cout << "Compute tiles: " << tile_n << "x" << tile_m << " per CTA, " << kSMs << " SMs (" << tiles
            << " tiles, " << waves << " waves)\n";
  std::cout << "Tiles in N: " << tiles_n << "\n";
  std::cout << "Tiles in M: " << tiles_m << "\n";
  const unsigned long long communicationVolume =
      static_cast<unsigned long long>(size_n) * size_m * (1 + size_k / tile_n + size_k / tile_m);
  std::cout << "Communication volume: " << communicationVolume << " elements ("
            << 1e-9 * communicationVolume * sizeof(Data_t) << " GB through L2 at this tile size)\n";
  const double ioAccesses = communicationVolume / (3 * static_cast<double>(size_n) * size_m * size_k);
  std::cout << "I/O access fraction: " << ioAccesses << "\n";
  const unsigned long long algorithmic =
      (static_cast<unsigned long long>(size_n) * size_k + static_cast<unsigned long long>(size_k) * size_m +
       static_cast<unsigned long long>(size_n) * size_m) * sizeof(Data_t);
  std::cout << "Algorithmic bytes:    " << algorithmic << "\n";
  return 0;
}
