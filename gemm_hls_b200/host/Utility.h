// Host utilities of the executables — counterpart of the reference's include/Utility.h.
//
// ReferenceImplementation() below is the SELF-CHECK the reference's executables run after the
// kernel (test/TestSimulation.cpp:61, host/RunHardware.cpp:198-227 `verify on`): a host triple
// loop with the configured functors.  It is never on the compute path: C always comes from
// libmm_b200.so, and with `verify off` this code does not run at all.  (The test-suite's oracle
// lives in oracle/ and is not linked here.)
#pragma once

#include <cstddef>
#include <iostream>
#include <limits>
#include <type_traits>
#include <vector>

#include "MatrixMultiplication.h"

namespace host_op {  // hlslib::op functors (hlslib/include/hlslib/xilinx/Operators.h:20-100), host side

template <typename T>
struct Add {
  static T Apply(T a, T b) { const T r = a + b; return r; }
  static T identity() { return T(0); }
};
template <typename T>
struct Multiply {
  static T Apply(T a, T b) { const T r = a * b; return r; }
  static T identity() { return T(1); }
};
template <typename T>
struct And {
  static T Apply(T a, T b) { return T((a != T(0)) && (b != T(0)) ? 1 : 0); }
  static T identity() { return T(1); }
};
template <typename T>
struct Min {
  static T Apply(T a, T b) { return (a < b) ? a : b; }
  static T identity() { return std::numeric_limits<T>::max(); }
};
template <typename T>
struct Max {
  static T Apply(T a, T b) { return (a > b) ? a : b; }
  static T identity() { return std::numeric_limits<T>::min(); }  // sic, Operators.h:96
};

template <typename T, int OP> struct Select;
template <typename T> struct Select<T, MM_OP_MULTIPLY> { using type = Multiply<T>; };
template <typename T> struct Select<T, MM_OP_ADD> { using type = Add<T>; };
template <typename T> struct Select<T, MM_OP_MIN> { using type = Min<T>; };
template <typename T> struct Select<T, MM_OP_MAX> { using type = Max<T>; };
template <typename T> struct Select<T, MM_OP_AND> { using type = And<T>; };

}  // namespace host_op

// half on the tensor cores (MM_HALF_TENSOR) accumulates in FP32 and rounds once: the host reference
// does the same; every other configuration computes in Data_t like the reference's Naive<>.
constexpr bool kHalfOnTensorCores = kDataIsHalf && !(kKernelFlags & MM_FLAG_EXACT) &&
                                    kMapOpCode == MM_OP_MULTIPLY && kReduceOpCode == MM_OP_ADD;
using Acc_t = std::conditional<kHalfOnTensorCores, float, Data_t>::type;
using OperatorMap = host_op::Select<Acc_t, kMapOpCode>::type;
using OperatorReduce = host_op::Select<Acc_t, kReduceOpCode>::type;

// include/Utility.h:105-111 (-> CallBLAS fallback :66-74 -> Naive :18-42)
inline void ReferenceImplementation(Data_t const *a, Data_t const *b, Data_t *c, const unsigned size_n,
                                    const unsigned size_k, const unsigned size_m) {
  for (unsigned n = 0; n < size_n; ++n) {
    for (unsigned m = 0; m < size_m; ++m) {
      Acc_t acc = OperatorReduce::identity();
      for (unsigned k = 0; k < size_k; ++k) {
#ifndef MM_TRANSPOSED_A
        const Data_t elem_a = a[static_cast<size_t>(n) * size_k + k];
#else
        const Data_t elem_a = a[static_cast<size_t>(k) * size_n + n];
#endif
        acc = OperatorReduce::Apply(acc, OperatorMap::Apply(Acc_t(elem_a), Acc_t(b[static_cast<size_t>(k) * size_m + m])));
      }
      c[static_cast<size_t>(n) * size_m + m] = Data_t(acc);
    }
  }
}

// include/Utility.h:113-129
template <typename T>
typename std::enable_if<std::is_integral<T>::value, signed long>::type make_signed(T val) {
  return static_cast<signed long>(val);
}
template <typename T>
typename std::enable_if<!std::is_integral<T>::value, T>::type make_signed(T val) {
  return val;
}

// The acceptance loop of test/TestSimulation.cpp:75-92 / host/RunHardware.cpp:207-224.
// Returns true when verified; prints the reference's mismatch line otherwise.
inline bool VerifyAgainstReference(std::vector<Data_t> const &test, std::vector<Data_t> const &ref,
                                   size_t size_n, size_t size_m) {
  for (size_t i = 0; i < size_n; ++i) {
    for (size_t j = 0; j < size_m; ++j) {
      const auto testVal = make_signed<Data_t>(test[i * size_m + j]);
      const auto refVal = make_signed<Data_t>(ref[i * size_m + j]);
      bool mismatch;
      if (std::is_integral<Data_t>::value) {
        mismatch = testVal != refVal;
      } else if (kDataIsHalf && !kHalfOnTensorCores) {
        // the reference's branch for half: std::is_floating_point<half> is false, so it compares
        // EXACTLY (test/TestSimulation.cpp:79-85) — the bit-exact datapath is held to that
        mismatch = !(static_cast<float>(testVal) == static_cast<float>(refVal));
      } else {
        // relative 1e-3 (the reference's floating-point branch), evaluated in double
        const double t = static_cast<double>(static_cast<float>(testVal));
        const double r = static_cast<double>(static_cast<float>(refVal));
        const double td = std::is_same<Data_t, double>::value ? static_cast<double>(testVal) : t;
        const double rd = std::is_same<Data_t, double>::value ? static_cast<double>(refVal) : r;
        const double diff = td > rd ? td - rd : rd - td;
        mismatch = diff / rd > 1e-3;
      }
      if (mismatch) {
        std::cerr << "Mismatch at (" << i << ", " << j << "): " << testVal << " vs. " << refVal << "\n";
        return false;
      }
    }
  }
  return true;
}
