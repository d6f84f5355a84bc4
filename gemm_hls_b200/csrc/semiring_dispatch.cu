// Run-time (dtype, map, reduce) -> compile-time instantiation of the semiring tile kernel.
#include "common.cuh"
#include "semiring_kernel.cuh"

namespace mm {

namespace {

template <typename T>
int by_map(int map_op, int reduce_op, const GemmArgs &g, bool ta, bool ring) {
  switch (map_op) {
    case MM_OP_MULTIPLY: return launch_semiring_for<T, MM_OP_MULTIPLY>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
    case MM_OP_ADD: return launch_semiring_for<T, MM_OP_ADD>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
    case MM_OP_MIN: return launch_semiring_for<T, MM_OP_MIN>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
    case MM_OP_MAX: return launch_semiring_for<T, MM_OP_MAX>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
    case MM_OP_AND: return launch_semiring_for<T, MM_OP_AND>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
  }
  return -1;
}

// float only: the hardware min/max variants (internal operator codes)
int by_map_float(int map_op, int reduce_op, const GemmArgs &g, bool ta, bool ring) {
  switch (map_op) {
    case MM_OP_MIN_FAST: return launch_semiring_for<float, MM_OP_MIN_FAST>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
    case MM_OP_MAX_FAST: return launch_semiring_for<float, MM_OP_MAX_FAST>(reduce_op, g.a, g.b, g.c, g.n, g.k, g.m, ta, ring, g.stream);
  }
  return by_map<float>(map_op, reduce_op, g, ta, ring);
}

}  // namespace

int launch_semiring(int dtype, int map_op, int reduce_op, const GemmArgs &g_in) {
  GemmArgs g = g_in;
  if (g.dry_run) g.a = nullptr;  // launch_semiring_typed: null A = load the kernel, launch nothing
  const bool ta = (g.flags & MM_FLAG_TRANSPOSED_A) != 0;
  const bool ring = g.tuning ? g.tuning->semiring_ring() : true;
  int rc = -1;
  switch (dtype) {
    case MM_DTYPE_HALF: rc = by_map<__half>(map_op, reduce_op, g, ta, ring); break;
    case MM_DTYPE_FLOAT: {
      // Min / Max on float use FMNMX unless the caller asked for the literal C++ semantics
      auto fast = [&](int op) {
        if (g.flags & MM_FLAG_EXACT) return op;
        return op == MM_OP_MIN ? int(MM_OP_MIN_FAST) : (op == MM_OP_MAX ? int(MM_OP_MAX_FAST) : op);
      };
      rc = by_map_float(fast(map_op), fast(reduce_op), g, ta, ring);
      break;
    }
    case MM_DTYPE_DOUBLE: rc = by_map<double>(map_op, reduce_op, g, ta, ring); break;
    case MM_DTYPE_INT32: rc = by_map<int>(map_op, reduce_op, g, ta, ring); break;
    case MM_DTYPE_UINT32: rc = by_map<unsigned>(map_op, reduce_op, g, ta, ring); break;
    case MM_DTYPE_UINT8: rc = by_map<unsigned char>(map_op, reduce_op, g, ta, ring); break;
    default: return fail(MM_ERR_INVALID, "unknown data type");
  }
  if (rc < 0) return fail(MM_ERR_INVALID, "unknown map/reduce operator");
  if (rc != 0) return fail(MM_ERR_CUDA, std::string("semiring kernel launch: ") + cudaGetErrorString(static_cast<cudaError_t>(rc)));
  return MM_OK;
}

}  // namespace mm
