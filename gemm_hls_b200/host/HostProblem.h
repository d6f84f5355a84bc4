// Shared plumbing of the three executables: command-line handling and the host-side problem
// instance (inputs drawn with the reference's recipe, the expected result for `verify on`, the
// acceptance check).  Behavioural contract = the reference's executables
// (test/TestSimulation.cpp, host/RunHardware.cpp); the code organisation is this project's own.
#pragma once

#include <cstddef>
#include <iostream>
#include <random>
#include <string>
#include <type_traits>
#include <vector>

#include "MatrixMultiplication.h"
#include "Utility.h"

namespace mmhost {

struct Shape {
  unsigned n = 0, k = 0, m = 0;
  size_t CountA() const { return static_cast<size_t>(n) * k; }
  size_t CountB() const { return static_cast<size_t>(k) * m; }
  size_t CountC() const { return static_cast<size_t>(n) * m; }
  double Operations() const { return 2.0 * static_cast<float>(n) * k * m; }  // host/RunHardware.cpp:174-176
};

// argv[first .. first+2] = N K M when sizes are dynamic; the compiled-in sizes otherwise.
// Returns the index of the next unread argument.
inline int ReadShape(char **argv, int first, Shape *shape) {
#ifdef MM_DYNAMIC_SIZES
  shape->n = static_cast<unsigned>(std::stoul(argv[first]));
  shape->k = static_cast<unsigned>(std::stoul(argv[first + 1]));
  shape->m = static_cast<unsigned>(std::stoul(argv[first + 2]));
  return first + 3;
#else
  (void)argv;
  shape->n = kSizeN;
  shape->k = kSizeK;
  shape->m = kSizeM;
  return first;
#endif
}

constexpr int kShapeArguments =
#ifdef MM_DYNAMIC_SIZES
    3;
#else
    0;
#endif

// The reference's divisibility rule (64-byte memory words along K and M).  `verbose` selects
// RunHardware's wording (host/RunHardware.cpp:50-61) over TestSimulation's (:22-35).
inline bool ShapeIsLegal(Shape const &s, bool verbose) {
  struct Dim {
    unsigned value;
    int width;
    const char *name;
  } dims[2] = {{s.k, kMemoryWidthK, "K"}, {s.m, kMemoryWidthM, "M"}};
  for (auto const &d : dims) {
    if (d.value % d.width == 0) continue;
    if (verbose) {
      std::cerr << d.name << " (" << d.value << ") must be divisable by the memory width in " << d.name << " ("
                << d.width << ")." << std::endl;
    } else {
      std::cerr << d.name << " must be divisable by memory width." << std::endl;
    }
    return false;
  }
  return true;
}

// Inputs, expected output and device result of one run.
class Problem {
 public:
  explicit Problem(Shape const &shape) : shape_(shape) {}

  // seed kSeed, U[1,10] reals (integers for integral Data_t), ALL of A first, then B
  // (test/TestSimulation.cpp:46-55)
  void DrawInputs() {
    a_.resize(shape_.CountA());
    b_.resize(shape_.CountB());
    std::default_random_engine engine(kSeed);
    using Distribution =
        typename std::conditional<std::is_integral<Data_t>::value, std::uniform_int_distribution<unsigned long>,
                                  std::uniform_real_distribution<double>>::type;
    Distribution draw(1, 10);
    for (auto *matrix : {&a_, &b_}) {
      for (auto &element : *matrix) element = Data_t(draw(engine));
    }
  }
  void AllocateOutputs() {
    expected_.assign(shape_.CountC(), Data_t(0));
    result_.assign(shape_.CountC(), Data_t(0));
  }
  void ComputeExpected() {
    ReferenceImplementation(a_.data(), b_.data(), expected_.data(), shape_.n, shape_.k, shape_.m);
  }
  bool ResultMatches() const { return VerifyAgainstReference(result_, expected_, shape_.n, shape_.m); }

  Data_t const *A() const { return a_.data(); }
  Data_t const *B() const { return b_.data(); }
  Data_t *Result() { return result_.data(); }
  Shape const &shape() const { return shape_; }

 private:
  Shape shape_;
  std::vector<Data_t> a_, b_, expected_, result_;
};

}  // namespace mmhost
