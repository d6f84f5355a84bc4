#!/usr/bin/env python
"""DMMA tile height (MM_DMMA_TILE_ROWS=128|64|auto) on the row blocks a G-GPU split of double 8192^3
hands to one GPU: rows = 8192 / G.  One process per setting (the env knob is read once)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
import gemm_hls_b200 as mm
rows = int(sys.argv[1]); k = m = 8192
ctx = mm.Context(0)
rng = np.random.default_rng(1)
a = rng.uniform(-1, 1, (rows, k)); b = rng.uniform(-1, 1, (k, m))
da, db, dc = ctx.alloc(a.nbytes), ctx.alloc(b.nbytes), ctx.alloc(rows * m * 8)
ctx.copy_to_device(da, a); ctx.copy_to_device(db, b)
ts = []
for i in range(8):
    dev, _ = ctx.execute(mm.DOUBLE, mm.MULTIPLY, mm.ADD, da, db, dc, rows, k, m)
    ts.append(dev)
c = np.empty((rows, m)); ctx.copy_to_host(c, dc)
ref = a[:4] @ b
err = float(np.max(np.abs(c[:4] - ref) / np.maximum(np.abs(ref), 1e-300)))
t = float(np.median(ts[3:]))
print(json.dumps({"rows": rows, "ms": t * 1e3, "tflops": 2.0 * rows * k * m / t / 1e12, "max_rel_err_4rows": err}))
'''

out = []
for rows in (8192, 4096, 2048, 1024):
    for setting in ("128", "64", "auto"):
        env = dict(os.environ)
        env.pop("MM_DMMA_TILE_ROWS", None)
        if setting != "auto":
            env["MM_DMMA_TILE_ROWS"] = setting
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, str(rows)], env=env, capture_output=True, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            print({"rows": rows, "tile_rows": setting, "error": "timeout"}, flush=True)
            continue
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        d = json.loads(line[-1]) if line else {"error": r.stderr[-400:]}
        d["tile_rows"] = setting
        out.append(d)
        print(d, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/exp_dmma_rows.json", "w"), indent=1)
