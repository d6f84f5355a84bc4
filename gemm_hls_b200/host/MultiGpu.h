// Row-block split of ONE MatrixMultiplicationKernel call over the GPUs of one box, native C++:
// one host thread per GPU, one NCCL communicator per GPU (ncclCommInitAll), B broadcast ONCE from
// GPU 0 over NVLink/NVSwitch, then every GPU runs the single-GPU C-ABI entry on its row-block.
// No per-step collective and no reduction: K is not split (outer tiles of C are independent in the
// reference, kernel/Compute.cpp:53-56).  Used by RunHardware.exe when MM_NUM_GPUS > 1.
#pragma once

#define CUDA_NO_HALF  // keep cuda_fp16.h from claiming the name `half` (HostTypes.h owns it, as in the reference)
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "mm_b200.h"

namespace mm {

class MultiGpuRun {
 public:
  MultiGpuRun(int gpus, int dtype, int map_op, int reduce_op, int flags, unsigned n, unsigned k, unsigned m,
              size_t elem)
      : gpus_(gpus), dtype_(dtype), map_(map_op), reduce_(reduce_op), flags_(flags), n_(n), k_(k), m_(m),
        elem_(elem) {
    if (flags & MM_FLAG_TRANSPOSED_A) throw std::runtime_error("row-block split needs row-major A");
    int count = 0;
    Cuda(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
    if (count < gpus) throw std::runtime_error("MM_NUM_GPUS exceeds the visible devices");
    ctx_.resize(gpus, nullptr);
    a_.resize(gpus, nullptr);
    b_.resize(gpus, nullptr);
    c_.resize(gpus, nullptr);
    comms_.resize(gpus);
    std::vector<int> devs(gpus);
    for (int g = 0; g < gpus; ++g) devs[g] = g;
    Nccl(ncclCommInitAll(comms_.data(), gpus, devs.data()), "ncclCommInitAll");
    const unsigned per = (n + gpus - 1) / gpus;  // ceil(N / G) rows per GPU, tail GPUs may get fewer
    for (int g = 0; g < gpus; ++g) {
      const unsigned r0 = std::min(n, g * per), r1 = std::min(n, r0 + per);
      rows_.push_back({r0, r1});
      Mm(mm_context_create(g, &ctx_[g]));
      Mm(mm_buffer_alloc(ctx_[g], size_t(std::max(1u, r1 - r0)) * k * elem, &a_[g]));
      Mm(mm_buffer_alloc(ctx_[g], size_t(k) * m * elem, &b_[g]));
      Mm(mm_buffer_alloc(ctx_[g], size_t(std::max(1u, r1 - r0)) * m * elem, &c_[g]));
    }
  }
  ~MultiGpuRun() {
    for (int g = 0; g < gpus_; ++g) {
      if (ctx_[g]) {
        mm_buffer_free(ctx_[g], a_[g]);
        mm_buffer_free(ctx_[g], b_[g]);
        mm_buffer_free(ctx_[g], c_[g]);
        mm_context_destroy(ctx_[g]);
      }
      ncclCommDestroy(comms_[g]);
    }
  }

  // A row-blocks to their GPUs, B to GPU 0 only.
  void CopyFromHost(const void *a, const void *b) {
    const unsigned char *pa = static_cast<const unsigned char *>(a);
    for (int g = 0; g < gpus_; ++g) {
      const size_t rows = rows_[g].second - rows_[g].first;
      if (rows) Mm(mm_copy_to_device(ctx_[g], a_[g], pa + size_t(rows_[g].first) * k_ * elem_, rows * k_ * elem_));
    }
    Mm(mm_copy_to_device(ctx_[0], b_[0], b, size_t(k_) * m_ * elem_));
  }

  // The path's one collective.  Returns wall seconds.
  double BroadcastB() {
    const auto t0 = std::chrono::high_resolution_clock::now();
    Nccl(ncclGroupStart(), "ncclGroupStart");
    for (int g = 0; g < gpus_; ++g) {
      Cuda(cudaSetDevice(g), "cudaSetDevice");
      Nccl(ncclBroadcast(b_[0 == g ? 0 : g], b_[g], size_t(k_) * m_ * elem_, ncclChar, 0, comms_[g], nullptr),
           "ncclBroadcast");
    }
    Nccl(ncclGroupEnd(), "ncclGroupEnd");
    for (int g = 0; g < gpus_; ++g) {
      Cuda(cudaSetDevice(g), "cudaSetDevice");
      Cuda(cudaStreamSynchronize(nullptr), "cudaStreamSynchronize");
    }
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  }

  // All GPUs multiply their row-block concurrently; {max device seconds, wall seconds}.
  std::pair<double, double> Execute() {
    std::vector<double> dev(gpus_, 0.0);
    std::vector<std::string> err(gpus_);
    std::vector<std::thread> pool;
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int g = 0; g < gpus_; ++g) {
      pool.emplace_back([&, g] {
        const unsigned rows = rows_[g].second - rows_[g].first;
        if (rows == 0) return;
        double d = 0, w = 0;
        if (mm_kernel_execute(ctx_[g], dtype_, map_, reduce_, flags_, a_[g], b_[g], c_[g], rows, k_, m_, &d, &w) !=
            MM_OK) {
          err[g] = mm_last_error();
        }
        dev[g] = d;
      });
    }
    for (auto &t : pool) t.join();
    for (auto &e : err) {
      if (!e.empty()) throw std::runtime_error(e);
    }
    const double wall = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    return {*std::max_element(dev.begin(), dev.end()), wall};
  }

  void CopyToHost(void *c) {
    unsigned char *pc = static_cast<unsigned char *>(c);
    for (int g = 0; g < gpus_; ++g) {
      const size_t rows = rows_[g].second - rows_[g].first;
      if (rows) Mm(mm_copy_to_host(ctx_[g], pc + size_t(rows_[g].first) * m_ * elem_, c_[g], rows * m_ * elem_));
    }
  }

 private:
  static void Mm(int rc) {
    if (rc != MM_OK) throw std::runtime_error(mm_last_error());
  }
  static void Cuda(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
  static void Nccl(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + ncclGetErrorString(r));
  }

  int gpus_, dtype_, map_, reduce_, flags_;
  unsigned n_, k_, m_;
  size_t elem_;
  std::vector<mm_context *> ctx_;
  std::vector<void *> a_, b_, c_;
  std::vector<ncclComm_t> comms_;
  std::vector<std::pair<unsigned, unsigned>> rows_;
};

}  // namespace mm
