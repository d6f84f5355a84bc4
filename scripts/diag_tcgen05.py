#!/usr/bin/env python
"""Device-resident spot checks of the tcgen05 variants at multi-tile sizes (diagnostics; the parity tests proper
are tests/): sampled rows of C against an FP64 torch.matmul, per tuning variant."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemm_hls_b200 as G

dev = torch.device("cuda", 0)
cases = [("half", 4096), ("half", 8192), ("half", 32768), ("float", 4096), ("float", 16384)]
variants = [dict(), dict(tma_store=0), dict(cta_group=1), dict(block_n=128), dict(tile_sync=0), dict(b_mn=0)]
if len(sys.argv) > 1:
    cases = [(c.split(":")[0], int(c.split(":")[1])) for c in sys.argv[1].split(",")]
for name, n in cases:
    tdt = torch.float16 if name == "half" else torch.float32
    dt = G.HALF if name == "half" else G.FLOAT
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    lo, hi = (0.0, 1.0) if name == "half" else (1.0, 10.0)
    a = (torch.rand((n, n), generator=gen, device=dev) * (hi - lo) + lo).to(tdt)
    b = (torch.rand((n, n), generator=gen, device=dev) * (hi - lo) + lo).to(tdt)
    rows = torch.tensor([0, 1, 127, 128, 255, 256, n // 2 + 3, n - 257, n - 1], device=dev)
    ref = a[rows].double() @ b.double()
    for v in variants:
        c = torch.full((n, n), float("nan"), device=dev, dtype=tdt)
        with G.Context(0) as ctx:
            ctx.set_tuning(**v)
            torch.cuda.synchronize()   # the context's stream is non-blocking: torch's fills must have finished
            ctx.enqueue(dt, G.MULTIPLY, G.ADD, a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n)
            torch.cuda.synchronize()
        got = c[rows].double()
        rel = ((got - ref).abs() / ref).max().item()
        bad = int((~torch.isfinite(c[::max(1, n // 64)])).sum().item())
        print(json.dumps({"type": name, "n": n, "variant": v, "max_rel": rel, "nonfinite_in_sample": bad}), flush=True)
    del a, b, ref
    torch.cuda.empty_cache()
