// RunHardware.exe N K M [hw|hw_emu] [on|off] — same argv grammar, stdout lines and exit codes as
// the reference's host/RunHardware.cpp:30-230 (its perf sentence is what
// scripts/build_manager.py:601-602 parses).  Context / MakeBuffer / CopyFromHost / MakeKernel /
// ExecuteTask / CopyToHost run against a B200 through the C-ABI (Device.h).
//   hw      -> the B200 (device 0, or $MM_DEVICE)
//   hw_emu  -> accepted for compatibility: there is no emulation target, it runs on the B200 too.
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <random>
#include <string>
#include <type_traits>
#include <vector>

#include "Device.h"
#include "MatrixMultiplication.h"
#include "Utility.h"
#ifdef MM_HAS_NCCL
#include "MultiGpu.h"
#endif

void PrintUsage() {
#ifndef MM_DYNAMIC_SIZES
  std::cerr << "Usage: ./RunHardware.exe <mode [hw/hw_emu]> [<verify [on/off]>]\n" << std::flush;
#else
  std::cerr << "Usage: ./RunHardware.exe N K M [<mode [hw/hw_emu]>] [<verify [on/off]>]\n" << std::flush;
#endif
}

int main(int argc, char **argv) {
  std::default_random_engine rng(kSeed);
  typename std::conditional<std::is_integral<Data_t>::value, std::uniform_int_distribution<unsigned long>,
                            std::uniform_real_distribution<double>>::type dist(1, 10);
  bool verify = true;
#ifdef MM_DYNAMIC_SIZES
  if (argc > 6 || argc < 4) {
    PrintUsage();
    return 1;
  }
  const unsigned size_n = std::stoul(argv[1]);
  const unsigned size_k = std::stoul(argv[2]);
  const unsigned size_m = std::stoul(argv[3]);
  int next_arg = 4;
  if (size_k % kMemoryWidthK != 0) {
    std::cerr << "K (" << size_k << ") must be divisable by the memory width in K (" << kMemoryWidthK << ")."
              << std::endl;
    return 1;
  }
  if (size_m % kMemoryWidthM != 0) {
    std::cerr << "M (" << size_m << ") must be divisable by the memory width in M (" << kMemoryWidthM << ")."
              << std::endl;
    return 1;
  }
#else
  if (argc > 3) {
    PrintUsage();
    return 1;
  }
  constexpr auto size_n = kSizeN;
  constexpr auto size_k = kSizeK;
  constexpr auto size_m = kSizeM;
  int next_arg = 1;
#endif
  if (next_arg < argc) {
    const std::string emulation_arg(argv[next_arg++]);
    if (emulation_arg != "hw" && emulation_arg != "hw_emu") {
      PrintUsage();
      return 1;
    }
  }
  if (next_arg < argc) {
    const std::string verify_arg(argv[next_arg++]);
    if (verify_arg == "off") {
      verify = false;
    } else if (verify_arg != "on") {
      PrintUsage();
      return 1;
    }
  }

  const size_t count_a = static_cast<size_t>(size_n) * size_k;
  const size_t count_b = static_cast<size_t>(size_k) * size_m;
  const size_t count_c = static_cast<size_t>(size_n) * size_m;
  std::vector<Data_t> a, b, cRef, cMem;
  std::cout << "Initializing host memory..." << std::flush;
  if (verify) {
    a = decltype(a)(count_a);
    std::for_each(a.begin(), a.end(), [&dist, &rng](Data_t &in) { in = Data_t(dist(rng)); });
    b = decltype(b)(count_b);
    std::for_each(b.begin(), b.end(), [&dist, &rng](Data_t &in) { in = Data_t(dist(rng)); });
    cRef = decltype(cRef)(count_c, Data_t(0));
    cMem = decltype(cMem)(count_c, Data_t(0));
  }
  std::cout << " Done.\n";

  // MM_NUM_GPUS=G (G > 1): split C row-blocks over G GPUs of this box, B broadcast once over NCCL
  // (SURVEY.md section 8e).  Same stdout contract; the reported time is the slowest GPU's kernel time.
  const char *gpus_env = std::getenv("MM_NUM_GPUS");
  const int num_gpus = gpus_env ? std::atoi(gpus_env) : 1;
  if (num_gpus > 1) {
#ifdef MM_HAS_NCCL
    try {
      std::cout << "Initializing " << num_gpus << " CUDA contexts and NCCL...\n" << std::flush;
      mm::MultiGpuRun run(num_gpus, kDataTypeCode, kMapOpCode, kReduceOpCode, kKernelFlags, size_n, size_k,
                          size_m, sizeof(Data_t));
      if (verify) {
        std::cout << "Copying memory to device...\n" << std::flush;
        run.CopyFromHost(a.data(), b.data());
      }
      std::cout << "Broadcasting B over NCCL...\n" << std::flush;
      const double bcast = run.BroadcastB();
      std::cout << "Executing kernel...\n" << std::flush;
      const auto elapsed = run.Execute();
      const auto perf = 1e-9 * (2 * static_cast<float>(size_n) * size_k * size_m) / elapsed.first;
      std::cout << "Kernel executed in " << elapsed.first << " seconds, corresponding to a performance of " << perf
                << " GOp/s.\n";
      std::cout << "NCCL broadcast of B took " << bcast << " seconds on " << num_gpus << " GPUs.\n";
      if (verify) {
        std::cout << "Copying back result...\n" << std::flush;
        run.CopyToHost(cMem.data());
      }
    } catch (std::runtime_error const &err) {
      std::cerr << "Execution failed with error: \"" << err.what() << "\"." << std::endl;
      return 1;
    }
#else
    std::cerr << "Execution failed with error: \"MM_NUM_GPUS > 1 needs a build with NCCL (MM_HAS_NCCL)\"." << std::endl;
    return 1;
#endif
  } else
  try {
    std::cout << "Initializing CUDA context...\n" << std::flush;
    const char *dev_env = std::getenv("MM_DEVICE");
    mm::Context context(dev_env ? std::atoi(dev_env) : 0);

    std::cout << "Initializing device memory...\n" << std::flush;
    auto aDevice = context.MakeBuffer<Data_t, mm::Access::read>(count_a);
    auto bDevice = context.MakeBuffer<Data_t, mm::Access::read>(count_b);
    auto cDevice = context.MakeBuffer<Data_t, mm::Access::write>(count_c);

    if (verify) {
      std::cout << "Copying memory to device...\n" << std::flush;
      aDevice.CopyFromHost(a.data());
      bDevice.CopyFromHost(b.data());
      cDevice.CopyFromHost(cMem.data());
    }
    // verify == off: like the reference (host/RunHardware.cpp:99-111,140-145) the device buffers
    // are left uninitialised; the timing does not depend on the values.

    std::cout << "Creating kernel...\n" << std::flush;
    auto kernel = context.MakeKernel(kDataTypeCode, kMapOpCode, kReduceOpCode, kKernelFlags, aDevice, bDevice,
                                     cDevice, size_n, size_k, size_m);

    std::cout << "Executing kernel...\n" << std::flush;
    const auto elapsed = kernel.ExecuteTask();

    const auto perf = 1e-9 * (2 * static_cast<float>(size_n) * size_k * size_m) / elapsed.first;
    std::cout << "Kernel executed in " << elapsed.first << " seconds, corresponding to a performance of " << perf
              << " GOp/s.\n";

    if (verify) {
      std::cout << "Copying back result...\n" << std::flush;
      cDevice.CopyToHost(cMem.data());
    }
  } catch (std::runtime_error const &err) {
    std::cerr << "Execution failed with error: \"" << err.what() << "\"." << std::endl;
    return 1;
  }

  if (verify) {
    std::cout << "Running reference implementation...\n" << std::flush;
    ReferenceImplementation(a.data(), b.data(), cRef.data(), size_n, size_k, size_m);
    std::cout << "Verifying result...\n" << std::flush;
    if (!VerifyAgainstReference(cMem, cRef, size_n, size_m)) return 1;
    std::cout << "Successfully verified." << std::endl;
  }
  return 0;
}
