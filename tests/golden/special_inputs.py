"""Deterministic non-recipe inputs shared by the golden-record generator (make_golden_special.py), the CPU oracle
tests and the GPU parity tests: mixed-sign data, NaN / signed zeros / infinities, full-range bytes.  numpy's
default_rng (PCG64) is bit-reproducible across machines for a given numpy version; the records carry the inputs'
SHA-256, so a generator drift shows up as a hash mismatch, not as a false parity failure."""
import hashlib

import numpy as np


def signed(np_dtype, n, k, m, seed, special=False):
    """Mixed-sign values without zeros; `special` sprinkles -0, +0, NaN and infinities (floating types).
    Integer types: small signed (or unsigned) integers."""
    rng = np.random.default_rng(seed)
    if np.issubdtype(np_dtype, np.floating):
        vals = np.array([-3.5, -1.25, -0.5, 0.75, 1.0, 2.5, 6.0], dtype=np.float64)
        a = rng.choice(vals, size=n * k)
        b = rng.choice(vals, size=k * m)
        if special:
            pool = np.array([-0.0, 0.0, np.nan, np.inf, -np.inf, -0.0, 0.0])
            for arr in (a, b):
                idx = rng.choice(arr.size, size=max(4, arr.size // 16), replace=False)
                arr[idx] = rng.choice(pool, size=idx.size)
        return a.astype(np_dtype), b.astype(np_dtype)
    lo = -4 if np.issubdtype(np_dtype, np.signedinteger) else 0
    return (rng.integers(lo, 5, size=n * k).astype(np_dtype), rng.integers(lo, 5, size=k * m).astype(np_dtype))


def full_range_bytes(n, k, m, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=n * k, dtype=np.uint8), rng.integers(0, 256, size=k * m, dtype=np.uint8)


def make(kind, np_dtype, n, k, m, seed):
    if kind == "bytes":
        return full_range_bytes(n, k, m, seed)
    return signed(np_dtype, n, k, m, seed, special=(kind == "special"))


def canonical_sha256(c):
    """SHA-256 of C with every NaN replaced by one canonical quiet NaN (the payload and sign of a NaN produced by
    inf - inf or 0 * inf are left open by the reference's C++); zeros keep their sign."""
    c = np.ascontiguousarray(c).copy()
    if np.issubdtype(c.dtype, np.floating):
        c[np.isnan(c)] = np.nan
    return hashlib.sha256(c.tobytes()).hexdigest()
