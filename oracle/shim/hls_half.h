// TEST INFRASTRUCTURE ONLY — stand-in for Xilinx <hls_half.h> (not vendored in
// the reference, absent from this image; used at include/Config.h.in:8-10 and
// include/Utility.h:125-129).  IEEE binary16 with round-to-nearest-even after
// EVERY operation, which is what a half-precision FPGA datapath (and the
// reference's Naive<> over `half`) computes.  Bit-level parity with Xilinx'
// own header is NOT pinned (see oracle/README.md: "half: parity unpinned").
#pragma once

#include <cmath>
#include <limits>
#include <ostream>

struct half {
  _Float16 v;
  half() : v(0) {}
  half(float f) : v(static_cast<_Float16>(f)) {}
  half(double d) : v(static_cast<_Float16>(d)) {}
  half(int i) : v(static_cast<_Float16>(i)) {}
  half(unsigned i) : v(static_cast<_Float16>(i)) {}
  half(long i) : v(static_cast<_Float16>(i)) {}
  half(unsigned long i) : v(static_cast<_Float16>(i)) {}
  operator float() const { return static_cast<float>(v); }
};

// binary16 x binary16 products are exact in binary32; sums of two binary16
// values are exact in binary32 unless the exponents differ by more than 13,
// where the single rounding below still gives the correctly rounded result
// because the small addend is then below half an ulp of the binary16 result.
inline half operator+(half a, half b) { return half(static_cast<float>(a) + static_cast<float>(b)); }
inline half operator-(half a, half b) { return half(static_cast<float>(a) - static_cast<float>(b)); }
inline half operator*(half a, half b) { return half(static_cast<float>(a) * static_cast<float>(b)); }
inline half operator/(half a, half b) { return half(static_cast<float>(a) / static_cast<float>(b)); }
inline bool operator<(half a, half b) { return static_cast<float>(a) < static_cast<float>(b); }
inline bool operator>(half a, half b) { return static_cast<float>(a) > static_cast<float>(b); }
inline bool operator==(half a, half b) { return static_cast<float>(a) == static_cast<float>(b); }
inline bool operator!=(half a, half b) { return static_cast<float>(a) != static_cast<float>(b); }
inline std::ostream &operator<<(std::ostream &os, half h) { return os << static_cast<float>(h); }

namespace std {
template <>
struct numeric_limits<half> {
  static constexpr bool is_specialized = true;
  static half max() { return half(65504.0f); }
  static half min() { return half(6.103515625e-05f); }
  static half lowest() { return half(-65504.0f); }
};
inline half abs(half h) { return half(std::fabs(static_cast<float>(h))); }
}  // namespace std
