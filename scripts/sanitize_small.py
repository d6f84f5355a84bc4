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

# (name, dtype, map, reduce, flags, (n, k, m)[, tuning])
CASES = [
    ("tcgen05_tf32", G.FLOAT, G.MULTIPLY, G.ADD, 0, (257, 48, 272)),
    ("tcgen05_tf32 multi-tile", G.FLOAT, G.MULTIPLY, G.ADD, 0, (600, 80, 528)),
    ("tcgen05_tf32 direct stores", G.FLOAT, G.MULTIPLY, G.ADD, 0, (257, 48, 272), dict(tma_store=0)),
    ("tcgen05_tf32 overlapped B", G.FLOAT, G.MULTIPLY, G.ADD, 0, (257, 48, 528), dict(b_overlap=1)),
    ("tcgen05_tf32 K-major B, 1 CTA, 128 cols", G.FLOAT, G.MULTIPLY, G.ADD, 0, (257, 48, 272), dict(b_mn=0, cta_group=1, block_n=128)),
    ("tcgen05_i8", G.UINT8, G.MULTIPLY, G.ADD, 0, (257, 192, 320)),
    ("tcgen05_i8 TA, 1 CTA", G.UINT8, G.MULTIPLY, G.ADD, G.FLAG_TRANSPOSED_A, (130, 64, 192), dict(cta_group=1)),
    ("tcgen05_tf32 TA", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_TRANSPOSED_A, (130, 64, 192)),
    ("tcgen05_tf32x3", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_TF32X3, (129, 48, 272)),
    ("tcgen05_f16", G.HALF, G.MULTIPLY, G.ADD, 0, (257, 96, 288)),
    ("dmma_f64", G.DOUBLE, G.MULTIPLY, G.ADD, 0, (130, 24, 136)),
    ("dmma_f64 TA", G.DOUBLE, G.MULTIPLY, G.ADD, G.FLAG_TRANSPOSED_A, (130, 24, 136)),
    ("dmma_f64 3 stages wrap", G.DOUBLE, G.MULTIPLY, G.ADD, 0, (70, 200, 264)),
    ("semiring f32 addmin", G.FLOAT, G.ADD, G.MIN, 0, (129, 48, 144)),
    ("semiring f32 exact", G.FLOAT, G.MULTIPLY, G.ADD, G.FLAG_EXACT, (129, 48, 144)),
    ("semiring i32", G.INT32, G.MULTIPLY, G.ADD, 0, (65, 32, 48)),
    ("semiring f32 addmin staged kernel", G.FLOAT, G.ADD, G.MIN, 0, (129, 48, 144), dict(semiring_ring=0)),
    ("semiring u8 exact", G.UINT8, G.MULTIPLY, G.ADD, G.FLAG_EXACT, (65, 128, 192)),
    ("semiring f16 exact", G.HALF, G.MULTIPLY, G.ADD, G.FLAG_EXACT, (65, 64, 96)),
    ("semiring f64 addmax TA", G.DOUBLE, G.ADD, G.MAX, G.FLAG_TRANSPOSED_A, (67, 16, 24)),
]
only = os.environ.get("SANITIZE_ONLY")  # substring filter on the case name, e.g. SANITIZE_ONLY=dmma
bad = 0
for case in CASES:
    name, dt, mp, rd, flags, (n, k, m) = case[:6]
    tuning = case[6] if len(case) > 6 else {}
    if only and only not in name:
        continue
    a, b = O.fill(dt, n, k, m, 3)
    if dt == G.HALF:
        a = (a.astype(np.float32) * np.float32(0.25)).astype(np.float16)
    with G.Context(0) as ctx:
        ctx.set_tuning(**tuning)
        c = ctx.gemm_host(dt, mp, rd, a, b, n, k, m, flags=flags)[0]
    ref = O.naive(dt, mp, rd, a, b, n, k, m, transposed_a=bool(flags & G.FLAG_TRANSPOSED_A), threads=4)
    ok = O.verify(dt, c, ref) == -1 if G.kernel_path(dt, mp, rd, flags) == "semiring_simt" or dt != G.HALF else True
    print("%-42s %s" % (name, "ok" if ok else "MISMATCH"), flush=True)
    bad += 0 if ok else 1
# the row-block split on one device listed twice: sliced upload of B, the gather kernel, host barriers
if not only or "multi" in only:
    for dt, shape in ((G.FLOAT, (300, 128, 272)), (G.HALF, (257, 128, 288)), (G.DOUBLE, (130, 128, 136))):
        n, k, m = shape
        a, b = O.fill(dt, n, k, m, 5)
        if dt == G.HALF:
            a = (a.astype(np.float32) * np.float32(0.25)).astype(np.float16)
        single = G.matrix_multiplication_kernel(a, b, n, k, m, dtype=dt)
        with G.Multi(2, devices=[0, 0]) as multi:
            c = multi.gemm_host(dt, G.MULTIPLY, G.ADD, a, b, n, k, m)[0]
        ok = c.tobytes() == single.tobytes()
        print("%-42s %s" % ("multi x2 dtype %d" % dt, "ok" if ok else "MISMATCH"), flush=True)
        bad += 0 if ok else 1
sys.exit(1 if bad else 0)
