#!/usr/bin/env python
"""DMMA kernel generations on double rows x 8192 x 8192: TMA-fed (MM_DMMA_TMA=1, default), LDGSTS producer
warps (MM_DMMA_TMA=0, MM_DMMA_PRODUCERS=1|4), every warp prefetching (MM_DMMA_WS=0).
These switches exist at commit ad15438 only; afterwards the TMA-fed kernel is the only one."""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
import gemm_hls_b200 as mm
k = m = 8192; n = int(os.environ['ROWS'])
ctx = mm.Context(0)
rng = np.random.default_rng(1)
a = rng.uniform(-1, 1, (n, k)); b = rng.uniform(-1, 1, (k, m))
da, db, dc = ctx.alloc(a.nbytes), ctx.alloc(b.nbytes), ctx.alloc(n * m * 8)
ctx.copy_to_device(da, a); ctx.copy_to_device(db, b)
ts = [ctx.execute(mm.DOUBLE, mm.MULTIPLY, mm.ADD, da, db, dc, n, k, m)[0] for _ in range(8)]
t = float(np.median(ts[3:]))
print(json.dumps({"ms": t * 1e3, "tflops": 2.0 * n * k * m / t / 1e12}))
'''
for rows in ("8192", "1024"):
    for tma, ws, pw in (("1", "1", "4"), ("0", "1", "4"), ("0", "1", "1"), ("0", "0", "1"), ("1", "1", "4")):
        env = dict(os.environ, MM_DMMA_TMA=tma, MM_DMMA_WS=ws, MM_DMMA_PRODUCERS=pw, ROWS=rows)
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print("rows=%s tma=%s ws=%s producers=%s" % (rows, tma, ws, pw), line[-1] if line else r.stderr[-300:], flush=True)
        except subprocess.TimeoutExpired:
            print("rows=%s tma=%s ws=%s producers=%s timeout" % (rows, tma, ws, pw), flush=True)
