#!/usr/bin/env python
"""Context numbers (GPU): what the vendor library (cuBLAS through torch.matmul) reaches on the same
shapes, so that the roofline fractions in profiles/ have a second, independent denominator.
Library calls are NOT on the product path; this script only measures."""
import json
import sys

import torch


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


out = {}
dev = torch.device("cuda")
for name, dt, n, tf32, iters in (("tf32_16384", torch.float32, 16384, True, 10), ("fp32_16384_no_tf32", torch.float32, 16384, False, 2),
                                 ("f16_16384", torch.float16, 16384, False, 10), ("bf16_8192", torch.bfloat16, 8192, False, 20),
                                 ("f64_8192", torch.float64, 8192, False, 3)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.rand((n, n), device=dev, dtype=torch.float32).to(dt)
    b = torch.rand((n, n), device=dev, dtype=torch.float32).to(dt)
    c = torch.empty((n, n), device=dev, dtype=dt)
    s = timeit(lambda: torch.matmul(a, b, out=c), iters)
    out[name] = {"seconds": s, "tflops": 2.0 * n ** 3 / s * 1e-12}
    del a, b, c
# HBM copy
x = torch.empty(1 << 30, device=dev, dtype=torch.float32)
y = torch.empty_like(x)
s = timeit(lambda: y.copy_(x), 10)
out["copy_4GiB_each_way"] = {"seconds": s, "gbs": 2 * x.numel() * 4 / s * 1e-9}
print(json.dumps(out, indent=1))
