// C++ RAII wrappers over the C-ABI with the CALL SHAPE of the reference's hlslib::ocl classes
// (hlslib/include/hlslib/common/OpenCL.h): Context (:366), Buffer (:502), Kernel (:1286) with
// MakeBuffer / CopyFromHost / CopyToHost / MakeKernel / ExecuteTask, so that RunHardware.cpp reads
// like the reference's host/RunHardware.cpp:116-190.  Errors become std::runtime_error, which the
// reference's main catches and reports (host/RunHardware.cpp:192-196).
#pragma once

#include <stdexcept>
#include <string>
#include <utility>

#include "mm_b200.h"

namespace mm {

class RuntimeError : public std::runtime_error {  // hlslib::ocl::RuntimeError, common/OpenCL.h:143-157
 public:
  using std::runtime_error::runtime_error;
};

inline void Check(int rc) {
  if (rc != MM_OK) throw RuntimeError(mm_last_error());
}

enum class Access { read, write, readWrite };  // common/OpenCL.h:119

class Context;

template <typename T, Access access>
class Buffer {
 public:
  Buffer() = default;
  Buffer(mm_context *ctx, size_t count) : ctx_(ctx), count_(count) {
    Check(mm_buffer_alloc(ctx_, count * sizeof(T), &ptr_));
  }
  Buffer(Buffer &&o) noexcept : ctx_(o.ctx_), ptr_(o.ptr_), count_(o.count_) { o.ptr_ = nullptr; }
  Buffer &operator=(Buffer &&o) noexcept {
    std::swap(ctx_, o.ctx_);
    std::swap(ptr_, o.ptr_);
    std::swap(count_, o.count_);
    return *this;
  }
  Buffer(Buffer const &) = delete;
  Buffer &operator=(Buffer const &) = delete;
  ~Buffer() {
    if (ptr_) mm_buffer_free(ctx_, ptr_);
  }
  // Buffer::CopyFromHost(iterator) / CopyToHost(iterator), common/OpenCL.h:648-720 (blocking)
  void CopyFromHost(T const *source) { Check(mm_copy_to_device(ctx_, ptr_, source, count_ * sizeof(T))); }
  void CopyToHost(T *target) const { Check(mm_copy_to_host(ctx_, target, ptr_, count_ * sizeof(T))); }
  void *devicePointer() const { return ptr_; }
  size_t nElements() const { return count_; }

 private:
  mm_context *ctx_ = nullptr;
  void *ptr_ = nullptr;
  size_t count_ = 0;
};

class Kernel {  // Program::MakeKernel(name, a, b, c, n, k, m) + Kernel::ExecuteTask()
 public:
  Kernel(mm_context *ctx, int dtype, int map_op, int reduce_op, int flags, void const *a, void const *b,
         void *c, unsigned n, unsigned k, unsigned m)
      : ctx_(ctx), dtype_(dtype), map_(map_op), reduce_(reduce_op), flags_(flags), a_(a), b_(b), c_(c),
        n_(n), k_(k), m_(m) {}
  // Returns {device seconds (CUDA events around the kernels), wall seconds}, like
  // Kernel::ExecuteTask()'s {CL_PROFILING END-START, chrono} (common/OpenCL.h:1486-1504).
  std::pair<double, double> ExecuteTask() {
    double dev = 0, wall = 0;
    Check(mm_kernel_execute(ctx_, dtype_, map_, reduce_, flags_, a_, b_, c_, n_, k_, m_, &dev, &wall));
    return {dev, wall};
  }

 private:
  mm_context *ctx_;
  int dtype_, map_, reduce_, flags_;
  void const *a_, *b_;
  void *c_;
  unsigned n_, k_, m_;
};

class Context {  // hlslib::ocl::Context, common/OpenCL.h:366-500
 public:
  explicit Context(int device = 0) { Check(mm_context_create(device, &ctx_)); }
  ~Context() { mm_context_destroy(ctx_); }
  Context(Context const &) = delete;
  Context &operator=(Context const &) = delete;

  template <typename T, Access access>
  Buffer<T, access> MakeBuffer(size_t count) {
    return Buffer<T, access>(ctx_, count);
  }
  template <typename TA, Access AA, typename TB, Access AB, typename TC, Access AC>
  Kernel MakeKernel(int dtype, int map_op, int reduce_op, int flags, Buffer<TA, AA> &a, Buffer<TB, AB> &b,
                    Buffer<TC, AC> &c, unsigned n, unsigned k, unsigned m) {
    return Kernel(ctx_, dtype, map_op, reduce_op, flags, a.devicePointer(), b.devicePointer(),
                  c.devicePointer(), n, k, m);
  }
  mm_context *handle() const { return ctx_; }

 private:
  mm_context *ctx_ = nullptr;
};

// The same lifecycle over G GPUs (mm_multi_*): C row-blocks, B uploaded in slices and assembled over NVLink.
class MultiContext {
 public:
  explicit MultiContext(int gpus) { Check(mm_multi_create(gpus, nullptr, &multi_)); }
  ~MultiContext() { mm_multi_destroy(multi_); }
  MultiContext(MultiContext const &) = delete;
  MultiContext &operator=(MultiContext const &) = delete;
  void Upload(int dtype, int flags, void const *a, void const *b, unsigned n, unsigned k, unsigned m) {
    Check(mm_multi_upload(multi_, dtype, flags, a, b, n, k, m));
  }
  std::pair<double, double> Execute(int dtype, int map_op, int reduce_op, int flags, unsigned n, unsigned k,
                                    unsigned m) {
    double dev = 0, wall = 0;
    Check(mm_multi_execute(multi_, dtype, map_op, reduce_op, flags, n, k, m, &dev, &wall));
    return {dev, wall};
  }
  void Download(int dtype, void *c, unsigned n, unsigned m) { Check(mm_multi_download(multi_, dtype, c, n, m)); }

 private:
  mm_multi *multi_ = nullptr;
};

}  // namespace mm
