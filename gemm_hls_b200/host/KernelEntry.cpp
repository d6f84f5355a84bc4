// extern "C" MatrixMultiplicationKernel for the configured (Data_t, OperatorMap, OperatorReduce):
// the drop-in for the reference's kernel/Top.cpp:9-117 entry when it is called directly with host
// pointers (test/TestSimulation.cpp:66).  All work happens in libmm_b200.so (mm_gemm_host).
#include <stdexcept>
#include <string>

#include "MatrixMultiplication.h"

extern "C" {

#ifdef MM_DYNAMIC_SIZES
void MatrixMultiplicationKernel(Data_t const a[], Data_t const b[], Data_t c[], const unsigned size_n,
                                const unsigned size_k, const unsigned size_m) {
#else
void MatrixMultiplicationKernel(Data_t const a[], Data_t const b[], Data_t c[]) {
  const unsigned size_n = kSizeN, size_k = kSizeK, size_m = kSizeM;
#endif
  const int rc = mm_gemm_host(nullptr, kDataTypeCode, kMapOpCode, kReduceOpCode, kKernelFlags, a, b, c,
                              size_n, size_k, size_m, nullptr, nullptr);
  if (rc != MM_OK) throw std::runtime_error(mm_last_error());
}
}
