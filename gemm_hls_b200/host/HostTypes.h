// Host-side stand-in for `half` (the reference uses Xilinx' hls_half.h, include/Config.h.in:8-10):
// IEEE binary16 storage with float conversion, rounding after every operation.
#pragma once

#include <cmath>
#include <limits>
#include <ostream>
#include <type_traits>

struct half {
  _Float16 v;
  half() : v(0) {}
  half(float f) : v(static_cast<_Float16>(f)) {}
  half(double d) : v(static_cast<_Float16>(d)) {}
  half(int i) : v(static_cast<_Float16>(i)) {}
  half(unsigned long i) : v(static_cast<_Float16>(i)) {}
  operator float() const { return static_cast<float>(v); }
};
inline half operator+(half a, half b) { return half(static_cast<float>(a) + static_cast<float>(b)); }
inline half operator-(half a, half b) { return half(static_cast<float>(a) - static_cast<float>(b)); }
inline half operator*(half a, half b) { return half(static_cast<float>(a) * static_cast<float>(b)); }
inline half operator/(half a, half b) { return half(static_cast<float>(a) / static_cast<float>(b)); }
inline bool operator<(half a, half b) { return static_cast<float>(a) < static_cast<float>(b); }
inline bool operator>(half a, half b) { return static_cast<float>(a) > static_cast<float>(b); }
inline std::ostream &operator<<(std::ostream &os, half h) { return os << static_cast<float>(h); }
namespace std {
template <>
struct numeric_limits<half> {
  static constexpr bool is_specialized = true;
  static half max() { return half(65504.0f); }
  static half min() { return half(6.103515625e-05f); }
};
}  // namespace std
