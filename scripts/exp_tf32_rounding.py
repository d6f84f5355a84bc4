#!/usr/bin/env python
"""Experiment (GPU): does tcgen05 kind::tf32 truncate or round fp32 operands?
Runs float 1024^3 (reference input recipe, all-positive U[1,10]) twice through the C-ABI:
  (a) production path: operands rounded to nearest TF32 by the prep kernels,
  (b) MM_EXPERIMENT_TF32_NO_ROUND=1: raw fp32 bits fed to the MMA,
and reports mean signed / max relative error against the oracle.  A mean signed error near
-2^-11 * 2 * 0.5 ~ -5e-4 in (b) means the hardware truncates.  Output: JSON on stdout."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np, gemm_hls_b200 as G, oracle as O
n = k = m = 1024
a, b = O.fill(O.FLOAT, n, k, m)
c = G.matrix_multiplication_kernel(a, b, n, k, m)
ref = O.naive(O.FLOAT, O.MULTIPLY, O.ADD, a, b, n, k, m, threads=16).astype(np.float64)
ex = a.reshape(n, k).astype(np.float64) @ b.reshape(k, m).astype(np.float64)
rel = (c.astype(np.float64) - ex) / ex
print(json.dumps({"mean_signed_rel_err_vs_fp64": float(rel.mean()), "max_rel_err_vs_fp64": float(np.abs(rel).max()),
                  "max_rel_err_vs_naive_float": float(np.abs((c - ref) / ref).max()),
                  "naive_float_max_rel_err_vs_fp64": float(np.abs((ref - ex) / ex).max())}))
''' % ROOT

out = {}
for label, env in (("rounded_rna", {}), ("raw_fp32_bits", {"MM_EXPERIMENT_TF32_NO_ROUND": "1"})):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=e)
    out[label] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]}
print(json.dumps(out, indent=1))
