// TestSimulation N K M — same invocation, messages and exit codes as the reference's
// test/TestSimulation.cpp:13-96; the "simulation" is the real B200 kernel behind the same
// extern "C" MatrixMultiplicationKernel(a, b, c, n, k, m) call with host pointers (:66).
#include <algorithm>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "MatrixMultiplication.h"
#include "Utility.h"

int main(int argc, char **argv) {
#ifdef MM_DYNAMIC_SIZES
  if (argc < 4 || argc > 4) {
    std::cerr << "Usage: ./TestSimulation N K M" << std::endl;
    return 1;
  }
  const unsigned size_n = std::stoul(argv[1]);
  const unsigned size_k = std::stoul(argv[2]);
  const unsigned size_m = std::stoul(argv[3]);
  if (size_k % kMemoryWidthK != 0) {
    std::cerr << "K must be divisable by memory width." << std::endl;
    return 1;
  }
  if (size_m % kMemoryWidthM != 0) {
    std::cerr << "M must be divisable by memory width." << std::endl;
    return 1;
  }
#else
  constexpr auto size_n = kSizeN;
  constexpr auto size_k = kSizeK;
  constexpr auto size_m = kSizeM;
#endif

  std::vector<Data_t> a(static_cast<size_t>(size_n) * size_k);
  std::vector<Data_t> b(static_cast<size_t>(size_k) * size_m);
  std::vector<Data_t> cReference(static_cast<size_t>(size_n) * size_m, Data_t(0));
  std::vector<Data_t> cTest(static_cast<size_t>(size_n) * size_m, Data_t(0));

  // the reference's input recipe, test/TestSimulation.cpp:46-55
  std::default_random_engine rng(kSeed);
  typename std::conditional<std::is_integral<Data_t>::value, std::uniform_int_distribution<unsigned long>,
                            std::uniform_real_distribution<double>>::type dist(1, 10);
  std::for_each(a.begin(), a.end(), [&dist, &rng](Data_t &in) { in = Data_t(dist(rng)); });
  std::for_each(b.begin(), b.end(), [&dist, &rng](Data_t &in) { in = Data_t(dist(rng)); });

  ReferenceImplementation(a.data(), b.data(), cReference.data(), size_n, size_k, size_m);

  std::cout << "Running simulation...\n" << std::flush;
  try {
#ifdef MM_DYNAMIC_SIZES
    MatrixMultiplicationKernel(a.data(), b.data(), cTest.data(), size_n, size_k, size_m);
#else
    MatrixMultiplicationKernel(a.data(), b.data(), cTest.data());
#endif
  } catch (std::runtime_error const &err) {
    std::cerr << "Execution failed with error: \"" << err.what() << "\"." << std::endl;
    return 1;
  }
  std::cout << "Verifying results...\n" << std::flush;
  if (!VerifyAgainstReference(cTest, cReference, size_n, size_m)) return 1;
  std::cout << "Matrix-matrix multiplication successfully verified.\n";
  return 0;
}
