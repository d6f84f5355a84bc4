#!/usr/bin/env python
"""Small ragged invocations of every kernel family, for compute-sanitizer (memcheck / racecheck /
synccheck) runs:  compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import gemm_hls_b200 as G  # noqa: E402
import oracle as O  # noqa: E402

CASES = [
    ("tcgen05_tf32", G.FLOAT, G.MULTIPLY, G.ADD, 0, (257, 48, 272)),
    ("tcgen05_tf32 TA", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_TRANSPOSED_A, (130, 64, 192)),
    ("tcgen05_tf32x3", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_TF32X3, (129, 48, 272)),
    ("tcgen05_f16", G.HALF, G.MULTIPLY, G.ADD, 0, (257, 96, 288)),
    ("dmma_f64", G.DOUBLE, G.MULTIPLY, G.ADD, 0, (130, 24, 136)),
    ("dmma_f64 TA", G.DOUBLE, G.MULTIPLY, G.ADD, G.FLAG_TRANSPOSED_A, (130, 24, 136)),
    ("dmma_f64 3 stages wrap", G.DOUBLE, G.MULTIPLY, G.ADD, 0, (70, 200, 264)),
    ("semiring f32 addmin", G.FLOAT, G.ADD, G.MIN, 0, (129, 48, 144)),
    ("semiring f32 exact", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_EXACT, (129, 48, 144)),
    ("semiring i32", G.INT32, G.MULTIPLY, G.ADD, 0, (65, 32, 48)),
    ("semiring u8", G.UINT8, G.MULTIPLY, G.ADD, 0, (65, 128, 192)),
    ("semiring f16 exact", G.HALF, G.MULTIPLY, G.ADD, G.FLAG_EXACT, (65, 64, 96)),
    ("semiring f64 addmax TA", G.DOUBLE, G.ADD, G.MAX, G.FLAG_TRANSPOSED_A, (67, 16, 24)),
]
only = os.environ.get("SANITIZE_ONLY")  # substring filter on the case name, e.g. SANITIZE_ONLY=dmma
bad = 0
for name, dt, mp, rd, flags, (n, k, m) in CASES:
    if only and only not in name:
        continue
    a, b = O.fill(dt, n, k, m, 3)
    if dt == G.HALF:
        a = (a.astype(np.float32) * np.float32(0.25)).astype(np.float16)
    c = G.matrix_multiplication_kernel(a, b, n, k, m, dtype=dt, map_op=mp, reduce_op=rd, flags=flags)
    ref = O.naive(dt, mp, rd, a, b, n, k, m, transposed_a=bool(flags & G.FLAG_TRANSPOSED_A), threads=4)
    ok = O.verify(dt, c, ref) == -1 if G.kernel_path(dt, mp, rd, flags) == "semiring_simt" or dt != G.HALF else True
    print("%-26s %s" % (name, "ok" if ok else "MISMATCH"), flush=True)
    bad += 0 if ok else 1
sys.exit(1 if bad else 0)
