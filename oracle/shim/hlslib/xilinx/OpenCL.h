// TEST INFRASTRUCTURE ONLY — placed EARLIER on the include path than the
// reference's hlslib/include so that include/Utility.h:13 resolves here instead
// of pulling <CL/cl2.hpp> (OpenCL/XRT are not in this image).  The only thing
// the CPU path of the reference needs from that header is
// hlslib::ocl::AlignedAllocator<T, 4096> (hlslib/common/OpenCL.h:1695-1790),
// used by Pack<> at include/Utility.h:44-54.
#pragma once

#include <cstddef>
#include <cstdlib>
#include <memory>
#include <new>
#include <type_traits>

namespace hlslib {
namespace ocl {

template <typename T, std::size_t alignment>
class AlignedAllocator {
 public:
  using value_type = T;
  using propagate_on_container_move_assignment = std::true_type;
  template <class U>
  struct rebind {
    using other = AlignedAllocator<U, alignment>;
  };
  AlignedAllocator() noexcept {}
  template <class U>
  AlignedAllocator(const AlignedAllocator<U, alignment> &) noexcept {}
  T *allocate(std::size_t n) {
    void *p = nullptr;
    if (posix_memalign(&p, alignment, n * sizeof(T) ? n * sizeof(T) : alignment) != 0) {
      throw std::bad_alloc();
    }
    return static_cast<T *>(p);
  }
  void deallocate(T *p, std::size_t) noexcept { std::free(p); }
  template <class U>
  bool operator==(const AlignedAllocator<U, alignment> &) const noexcept { return true; }
  template <class U>
  bool operator!=(const AlignedAllocator<U, alignment> &) const noexcept { return false; }
};

}  // namespace ocl
}  // namespace hlslib
