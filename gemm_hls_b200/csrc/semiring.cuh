// Semiring functors for the device kernels: same interface as the reference's hlslib::op
// functors (hlslib/include/hlslib/xilinx/Operators.h:20-100) — `static T Apply(a, b)` and
// `static constexpr T identity()` — so a (Map, Reduce) pair plugs into the kernels exactly like
// MM_MAP_OP / MM_REDUCE_OP plug into the reference's ProcessingElement (kernel/Compute.cpp:129-133).
//
// Rounding contract: Apply() performs ONE correctly rounded operation in T (no FMA contraction
// across Map and Reduce), so a kernel that reduces sequentially over k reproduces the
// reference's Naive<> (include/Utility.h:18-42) bit for bit.
#pragma once

#include <cuda_fp16.h>

#include <cfloat>
#include <climits>
#include <cstdint>

#include "../../include/mm_b200.h"

namespace mm {

// ---- per-type primitives ---------------------------------------------------------------------
template <typename T>
struct Prim {
  static __host__ __device__ __forceinline__ T add(T a, T b) { return static_cast<T>(a + b); }
  static __host__ __device__ __forceinline__ T mul(T a, T b) { return static_cast<T>(a * b); }
  static __host__ __device__ __forceinline__ bool lt(T a, T b) { return a < b; }
  static __host__ __device__ __forceinline__ bool nz(T a) { return a != T(0); }
  static __host__ __device__ __forceinline__ T zero() { return T(0); }
  static __host__ __device__ __forceinline__ T one() { return T(1); }
};

template <>
struct Prim<float> {
  // __fadd_rn/__fmul_rn are never contracted into an FMA by the compiler.
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ bool lt(float a, float b) { return a < b; }
  static __device__ __forceinline__ bool nz(float a) { return a != 0.0f; }
  static __device__ __forceinline__ float zero() { return 0.0f; }
  static __device__ __forceinline__ float one() { return 1.0f; }
  // numeric_limits<float>::max() / ::min()
  static __device__ __forceinline__ float max_value() { return FLT_MAX; }
  static __device__ __forceinline__ float min_value() { return FLT_MIN; }
};

template <>
struct Prim<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ bool lt(double a, double b) { return a < b; }
  static __device__ __forceinline__ bool nz(double a) { return a != 0.0; }
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double one() { return 1.0; }
  static __device__ __forceinline__ double max_value() { return DBL_MAX; }
  static __device__ __forceinline__ double min_value() { return DBL_MIN; }
};

template <>
struct Prim<__half> {
  static __device__ __forceinline__ __half add(__half a, __half b) { return __hadd_rn(a, b); }
  static __device__ __forceinline__ __half mul(__half a, __half b) { return __hmul_rn(a, b); }
  static __device__ __forceinline__ bool lt(__half a, __half b) { return __hlt(a, b); }
  static __device__ __forceinline__ bool nz(__half a) { return __hneu(a, __ushort_as_half(0)); }  // `a != 0`: true for NaN
  static __device__ __forceinline__ __half zero() { return __ushort_as_half(0x0000); }
  static __device__ __forceinline__ __half one() { return __ushort_as_half(0x3C00); }
  static __device__ __forceinline__ __half max_value() { return __ushort_as_half(0x7BFF); }  // 65504
  static __device__ __forceinline__ __half min_value() { return __ushort_as_half(0x0400); }  // 2^-14
};

template <typename T>
struct IntLimits;
template <>
struct IntLimits<int> {
  static __host__ __device__ __forceinline__ int max_value() { return INT_MAX; }
  static __host__ __device__ __forceinline__ int min_value() { return INT_MIN; }
};
template <>
struct IntLimits<unsigned> {
  static __host__ __device__ __forceinline__ unsigned max_value() { return UINT_MAX; }
  static __host__ __device__ __forceinline__ unsigned min_value() { return 0u; }
};
template <>
struct IntLimits<unsigned char> {
  static __host__ __device__ __forceinline__ unsigned char max_value() { return 255; }
  static __host__ __device__ __forceinline__ unsigned char min_value() { return 0; }
};

template <typename T>
struct Lim {
  static __device__ __forceinline__ T max_value() { return IntLimits<T>::max_value(); }
  static __device__ __forceinline__ T min_value() { return IntLimits<T>::min_value(); }
};
template <> struct Lim<float> : Prim<float> {};
template <> struct Lim<double> : Prim<double> {};
template <> struct Lim<__half> : Prim<__half> {};

// ---- the functors (Operators.h) ---------------------------------------------------------------
template <typename T>
struct Sum {  // Operators.h:20-33
  static __device__ __forceinline__ T Apply(T a, T b) { return Prim<T>::add(a, b); }
  static __device__ __forceinline__ T identity() { return Prim<T>::zero(); }
};
template <typename T>
using Add = Sum<T>;  // Operators.h:35-36

template <typename T>
struct Product {  // Operators.h:45-58
  static __device__ __forceinline__ T Apply(T a, T b) { return Prim<T>::mul(a, b); }
  static __device__ __forceinline__ T identity() { return Prim<T>::one(); }
};
template <typename T>
using Multiply = Product<T>;  // Operators.h:60-61

template <typename T>
struct And {  // Operators.h:63-74 — `a && b` converted back to T
  static __device__ __forceinline__ T Apply(T a, T b) {
    return (Prim<T>::nz(a) && Prim<T>::nz(b)) ? Prim<T>::one() : Prim<T>::zero();
  }
  static __device__ __forceinline__ T identity() { return Prim<T>::one(); }
};

template <typename T>
struct Min {  // Operators.h:76-87
  static __device__ __forceinline__ T Apply(T a, T b) { return Prim<T>::lt(a, b) ? a : b; }
  static __device__ __forceinline__ T identity() { return Lim<T>::max_value(); }
};

template <typename T>
struct Max {  // Operators.h:89-100 — identity is numeric_limits<T>::min() as in the reference
  static __device__ __forceinline__ T Apply(T a, T b) { return Prim<T>::lt(b, a) ? a : b; }
  static __device__ __forceinline__ T identity() { return Lim<T>::min_value(); }
};

// Hardware min / max for float (FMNMX, and FMNMX3 when the compiler fuses two reductions): one
// instruction instead of the compare + select that the literal `(a < b) ? a : b` needs.  Identical
// results for all finite inputs except the sign of a zero when the operands are -0 and +0, and
// NaNs are dropped instead of propagated; selected for float unless MM_FLAG_EXACT is given.
template <typename T>
struct MinFast : Min<T> {};
template <typename T>
struct MaxFast : Max<T> {};
template <>
struct MinFast<float> {
  static __device__ __forceinline__ float Apply(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float identity() { return Lim<float>::max_value(); }
};
template <>
struct MaxFast<float> {
  static __device__ __forceinline__ float Apply(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float identity() { return Lim<float>::min_value(); }
};

// ---- packed pairs (float) ------------------------------------------------------------------------
// Blackwell issues two IEEE round-to-nearest FP32 additions / multiplications in ONE instruction
// (add.rn.f32x2 / mul.rn.f32x2 -> FADD2 / FMUL2; a scalar operand is broadcast for free), each half
// rounded exactly like __fadd_rn / __fmul_rn.  The tile kernel uses them for two adjacent columns
// of C at a time, which halves the issue slots of a float Map (and of a float Sum / Product Reduce)
// without touching the per-element order of operations.
struct F32x2 {
  unsigned long long v;
};
__device__ __forceinline__ F32x2 pack_f32x2(float lo, float hi) {
  F32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(F32x2 p, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p.v));
}
template <class Op>
struct PackedOp {
  static constexpr bool value = false;
};
template <>
struct PackedOp<Sum<float>> {
  static constexpr bool value = true;
  static __device__ __forceinline__ F32x2 Apply2(F32x2 a, F32x2 b) {
    F32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
  }
};
template <>
struct PackedOp<Product<float>> {
  static constexpr bool value = true;
  static __device__ __forceinline__ F32x2 Apply2(F32x2 a, F32x2 b) {
    F32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
  }
};

// ---- packed pairs (half) -------------------------------------------------------------------------
// HADD2 / HMUL2 do two IEEE round-to-nearest half operations per instruction, each half rounded exactly like the
// scalar __hadd_rn / __hmul_rn; the _rn intrinsics are never contracted into an HFMA2 (a single rounding, which
// Naive<> does not do).  Used when BOTH Map and Reduce are Sum / Product — (Multiply, Add) under MM_FLAG_EXACT, the
// datapath the half host programs run by default — for two adjacent columns of C at a time.
template <class Op>
struct PackedOpH {
  static constexpr bool value = false;
};
template <>
struct PackedOpH<Sum<__half>> {
  static constexpr bool value = true;
  static __device__ __forceinline__ __half2 Apply2(__half2 a, __half2 b) { return __hadd2_rn(a, b); }
};
template <>
struct PackedOpH<Product<__half>> {
  static constexpr bool value = true;
  static __device__ __forceinline__ __half2 Apply2(__half2 a, __half2 b) { return __hmul2_rn(a, b); }
};

// internal operator codes (never cross the C-ABI)
enum { MM_OP_MIN_FAST = 5, MM_OP_MAX_FAST = 6 };

template <typename T, int OP>
struct OpSelect;
template <typename T> struct OpSelect<T, MM_OP_MULTIPLY> { using type = Product<T>; };
template <typename T> struct OpSelect<T, MM_OP_ADD> { using type = Sum<T>; };
template <typename T> struct OpSelect<T, MM_OP_MIN> { using type = Min<T>; };
template <typename T> struct OpSelect<T, MM_OP_MAX> { using type = Max<T>; };
template <typename T> struct OpSelect<T, MM_OP_AND> { using type = And<T>; };
template <typename T> struct OpSelect<T, MM_OP_MIN_FAST> { using type = MinFast<T>; };
template <typename T> struct OpSelect<T, MM_OP_MAX_FAST> { using type = MaxFast<T>; };

// MM_DTYPE_* -> C type
template <int DTYPE> struct DTypeOf;
template <> struct DTypeOf<MM_DTYPE_HALF> { using type = __half; };
template <> struct DTypeOf<MM_DTYPE_FLOAT> { using type = float; };
template <> struct DTypeOf<MM_DTYPE_DOUBLE> { using type = double; };
template <> struct DTypeOf<MM_DTYPE_INT32> { using type = int; };
template <> struct DTypeOf<MM_DTYPE_UINT32> { using type = unsigned; };
template <> struct DTypeOf<MM_DTYPE_UINT8> { using type = unsigned char; };

}  // namespace mm
