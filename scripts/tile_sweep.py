#!/usr/bin/env python
"""Tile sweep of the tensor-core path (GPU) — the B200 counterpart of the reference's
scripts/build_manager.py `benchmark` (:578-669): run each configuration, record `config, time,
performance` in a benchmark.csv, here together with the DRAM traffic ncu measures, next to the
reference's own communication-volume model  Q = N*M*(1 + K/T_N + K/T_M)  elements
(src/PrintSpecifications.cpp:72-78) evaluated for the PATCH of C that co-running tiles share through
L2 (T_N = rasterisation-group rows, T_M = co-running column extent).

    python scripts/tile_sweep.py --workload half32768 --out gpurun_out/r02_tile_sweep_half32768.csv

Where the reference rebuilds the bitstream per configuration (MM_PARALLELISM_*, MM_MEMORY_TILE_SIZE_*,
scripts/build_manager.py:224-306), every variant here is compiled into libmm_b200.so and selected per context
(mm_context_set_tuning; bench.py --tune).  Swept: CTA group (1 | 2 = cta_group::2 pairs), C tile columns
(UMMA N = 128 | 256), ring depth, rasterisation-group rows, epilogue route (TMA stores | direct), B layout
(MN-major in place | K-major copy).  Every configuration passes bench.py's on-device result check before its
line is written (column `verified`); the parity suite proper runs the same knobs (tests/test_variants_gpu.py).
"""
import argparse
import csv
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {"half32768": (32768, 2), "float16384": (16384, 4), "half16384": (16384, 2)}


def tune_arg(cfg):
    return ",".join("%s=%d" % kv for kv in sorted(cfg.items()))


def run_bench(workload, cfg, steps):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", str(steps),
                        "--warmup", "3", "--no-e2e", "--no-cpu", "--tune", tune_arg(cfg)], capture_output=True, text=True,
                       timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if (lines and r.returncode == 0) else None   # non-zero: the result check failed


def run_ncu(workload, cfg):
    e = dict(os.environ)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct",
           "--clock-control", "none", "-k", "regex:gemm_tcgen05", "-s", "1", "-c", "1", "--csv",
           sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "1", "--warmup", "3",
           "--no-e2e", "--no-cpu", "--tune", tune_arg(cfg)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=900)
    vals = {}
    for row in csv.reader(l for l in r.stdout.splitlines() if l.startswith('"')):
        if len(row) > 3 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"):
            v = float(row[-1].replace(",", ""))
            unit = row[-2]
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
            vals[row[-3]] = v
    return vals


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="half32768")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tile_sweep.csv"))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--small", action="store_true", help="nine configurations around the default (one knob at a time)")
    ap.add_argument("--no-ncu", action="store_true", help="skip the DRAM-traffic capture (faster)")
    args = ap.parse_args()
    size, eb = SHAPES.get(args.workload, (16384, 4))
    base = dict(cta_group=2, block_n=256, stages=0, raster_rows=2048, tma_store=1, b_mn=1)
    if args.quick:
        grid = [dict(base), dict(base, cta_group=1)]
    elif args.small:
        grid = [dict(base), dict(base, cta_group=1), dict(base, block_n=128), dict(base, stages=3), dict(base, stages=4),
                dict(base, raster_rows=1024), dict(base, raster_rows=4096), dict(base, tma_store=0), dict(base, b_mn=0)]
    else:
        grid = []
        for cg, bn in ((2, 256), (1, 256), (2, 128), (1, 128)):
            deepest = {(2, 256): 6, (1, 256): 4, (2, 128): 8, (1, 128): 6}[(cg, bn)]
            for st in sorted({3, 4, deepest}):
                if st <= deepest:
                    grid.append(dict(base, cta_group=cg, block_n=bn, stages=st))
        grid += [dict(base, raster_rows=rr) for rr in (256, 512, 1024, 4096, 8192)]
        grid += [dict(base, tma_store=0), dict(base, b_mn=0), dict(base, tile_sync=0)]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["config", "time", "performance", "cta_group", "block_n", "stages", "raster_rows", "tma_store", "b_mn",
                    "verified", "dram_bytes", "model_bytes", "l2_hit_pct", "sm_mhz"])
        for cfg in grid:
            cg, bn, rr = cfg["cta_group"], cfg["block_n"], cfg["raster_rows"]
            d = run_bench(args.workload, cfg, args.steps)
            name = "%s_%s" % (args.workload, tune_arg({k: v for k, v in cfg.items() if base.get(k) != v}) or "default")
            if d is None:
                w.writerow([name, "", "", cg, bn, cfg["stages"], rr, cfg["tma_store"], cfg["b_mn"], "FAILED", "", "", "", ""])
                print(name, "FAILED", flush=True)
                continue
            t = run_ncu(args.workload, cfg) if not args.no_ncu else {}
            # patch shared through L2: rr rows x (co-running tiles / row-tiles-per-group) column tiles
            tile_rows = 128 * cg
            groups = 148 // cg
            rows_tiles = max(1, min(rr, size) // tile_rows)
            t_n = rows_tiles * tile_rows
            t_m = max(float(bn), groups / rows_tiles * float(bn))
            model = eb * size * size * (1 + size / t_n + size / min(t_m, size))
            r = d["roofline"]
            dram = t.get("dram__bytes_read.sum", 0) + t.get("dram__bytes_write.sum", 0)
            w.writerow([name, "%.6f" % (1e-3 * r["kernel_ms"]), "%.1f" % (1e3 * r["achieved"]), cg, bn, cfg["stages"], rr,
                        cfg["tma_store"], cfg["b_mn"], "yes" if d.get("check") else "n/a", int(dram), int(model),
                        "%.1f" % t.get("lts__t_sector_hit_rate.pct", float("nan")), d["clocks"]["sm_mhz"]])
            f.flush()
            print(name, "%.3f ms" % r["kernel_ms"], "%.0f GOp/s" % (1e3 * r["achieved"]),
                  "dram %.1f GB (model %.1f GB)" % (dram * 1e-9, model * 1e-9), flush=True)


if __name__ == "__main__":
    main()
