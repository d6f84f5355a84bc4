"""ORACLE — test infrastructure only (see oracle/naive.cpp, oracle/README.md).

ctypes bindings over oracle/liboracle.so (my restatement of the reference's Naive<> + input
recipe + acceptance criterion) and oracle/_ref/libref_naive_<cfg>.so (the reference's own
Naive<>, include/Utility.h:18-42, compiled in place by oracle/build.py).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package gemm_hls_b200 never imports this module.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# dtype / op codes mirror include/mm_b200.h
HALF, FLOAT, DOUBLE, INT32, UINT32, UINT8 = range(6)
MULTIPLY, ADD, MIN, MAX, AND = range(5)
NP_DTYPE = {HALF: np.float16, FLOAT: np.float32, DOUBLE: np.float64,
            INT32: np.int32, UINT32: np.uint32, UINT8: np.uint8}
DTYPE_NAME = {HALF: "half", FLOAT: "float", DOUBLE: "double", INT32: "int",
              UINT32: "unsigned", UINT8: "uint8_t"}
OP_NAME = {MULTIPLY: "Multiply", ADD: "Add", MIN: "Min", MAX: "Max", AND: "And"}

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            from oracle import build
            build.build_oracle()
        _lib = ctypes.CDLL(path)
        _lib.oracle_naive_rows.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3 + \
            [ctypes.c_long] * 5 + [ctypes.c_int]
        _lib.oracle_naive_rows.restype = ctypes.c_int
        _lib.oracle_fill.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
        _lib.oracle_fill.restype = ctypes.c_int
        _lib.oracle_verify.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        _lib.oracle_verify.restype = ctypes.c_long
    return _lib


def fill(dtype, n, k, m, seed=5):
    """Reference input recipe (test/TestSimulation.cpp:42-55): A (n*k) drawn first, then B (k*m)."""
    a = np.empty(n * k, dtype=NP_DTYPE[dtype])
    b = np.empty(k * m, dtype=NP_DTYPE[dtype])
    rc = lib().oracle_fill(dtype, a.ctypes.data, a.size, b.ctypes.data, b.size, seed)
    assert rc == 0
    return a, b


def naive(dtype, map_op, reduce_op, a, b, n, k, m, transposed_a=False, rows=None, threads=1):
    """C = A (x) B per include/Utility.h:18-42.  rows=(r0, r1) computes only those rows (others 0)."""
    a = np.ascontiguousarray(a, dtype=NP_DTYPE[dtype]).reshape(-1)
    b = np.ascontiguousarray(b, dtype=NP_DTYPE[dtype]).reshape(-1)
    assert a.size == n * k and b.size == k * m
    c = np.zeros(n * m, dtype=NP_DTYPE[dtype])
    r0, r1 = (0, n) if rows is None else rows
    rc = lib().oracle_naive_rows(dtype, map_op, reduce_op, int(transposed_a), a.ctypes.data,
                                 b.ctypes.data, c.ctypes.data, n, k, m, r0, r1, threads)
    if rc != 0:
        raise ValueError("oracle_naive_rows failed with code %d" % rc)
    return c.reshape(n, m)


def verify(dtype, test, ref):
    """Reference acceptance criterion (test/TestSimulation.cpp:75-92): flat index of first mismatch or -1."""
    test = np.ascontiguousarray(test, dtype=NP_DTYPE[dtype]).reshape(-1)
    ref = np.ascontiguousarray(ref, dtype=NP_DTYPE[dtype]).reshape(-1)
    assert test.size == ref.size
    return int(lib().oracle_verify(dtype, test.ctypes.data, ref.ctypes.data, test.size))


def ref_config_name(dtype, map_op, reduce_op, transposed_a=False):
    return "%s_%s_%s%s" % (DTYPE_NAME[dtype], OP_NAME[map_op], OP_NAME[reduce_op],
                           "_TA" if transposed_a else "")


def ref_available(dtype, map_op, reduce_op, transposed_a=False):
    return os.path.exists(os.path.join(
        HERE, "_ref", "libref_naive_%s.so" % ref_config_name(dtype, map_op, reduce_op, transposed_a)))


_ref_libs = {}


def ref_lib(dtype, map_op, reduce_op, transposed_a=False):
    name = ref_config_name(dtype, map_op, reduce_op, transposed_a)
    if name not in _ref_libs:
        l = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_naive_%s.so" % name))
        l.ref_naive.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3
        l.ref_naive.restype = None
        assert l.ref_sizeof_data_t() == np.dtype(NP_DTYPE[dtype]).itemsize
        _ref_libs[name] = l
    return _ref_libs[name]


def ref_naive(dtype, map_op, reduce_op, a, b, n, k, m, transposed_a=False):
    """The reference's OWN Naive<> (compiled in place into oracle/_ref/); single-threaded as written."""
    l = ref_lib(dtype, map_op, reduce_op, transposed_a)
    a = np.ascontiguousarray(a, dtype=NP_DTYPE[dtype]).reshape(-1)
    b = np.ascontiguousarray(b, dtype=NP_DTYPE[dtype]).reshape(-1)
    c = np.zeros(n * m, dtype=NP_DTYPE[dtype])
    l.ref_naive(a.ctypes.data, b.ctypes.data, c.ctypes.data, n, k, m)
    return c.reshape(n, m)
