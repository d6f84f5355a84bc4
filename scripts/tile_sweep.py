#!/usr/bin/env python
"""Tile sweep of the tensor-core path (GPU) — the B200 counterpart of the reference's
scripts/build_manager.py `benchmark` (:578-669): run each configuration, record `config, time,
performance` in a benchmark.csv, here together with the DRAM traffic ncu measures, next to the
reference's own communication-volume model  Q = N*M*(1 + K/T_N + K/T_M)  elements
(src/PrintSpecifications.cpp:72-78) evaluated for the PATCH of C that co-running tiles share through
L2 (T_N = rasterisation-group rows, T_M = co-running column extent).

    python scripts/tile_sweep.py --workload half32768 --out gpurun_out/r01/tile_sweep_half32768.csv

Swept: CTA group (1 | 2 = cta_group::2 pairs), ring depth, rasterisation-group rows.
Each configuration runs in its own process (the knobs are environment variables read once).
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


def run_bench(workload, env, steps):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", str(steps),
                        "--warmup", "3", "--no-e2e", "--no-cpu"], capture_output=True, text=True, env=e, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def run_ncu(workload, env):
    e = dict(os.environ)
    e.update(env)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct",
           "--clock-control", "none", "-k", "regex:gemm_tcgen05", "-s", "1", "-c", "1", "--csv",
           sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "1", "--warmup", "3",
           "--no-e2e", "--no-cpu"]
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
    args = ap.parse_args()
    size, eb = SHAPES.get(args.workload, (16384, 4))
    if args.quick:
        grid = [(2, 4, 2048), (1, 4, 2048)]
    else:
        grid = [(cg, st, rr) for cg, st, rr in itertools.product((1, 2), (3, 4, 5, 6), (2048,)) if not (cg == 1 and st > 4)]
        grid += [(2, 4, rr) for rr in (256, 512, 1024, 4096, 8192)]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["config", "time", "performance", "cta_group", "stages", "raster_rows", "dram_bytes",
                    "model_bytes", "l2_hit_pct", "sm_mhz"])
        for cg, st, rr in grid:
            env = {"MM_TCGEN05_CTA_GROUP": str(cg), "MM_TCGEN05_STAGES": str(st), "MM_TCGEN05_RASTER_ROWS": str(rr)}
            d = run_bench(args.workload, env, args.steps)
            t = run_ncu(args.workload, env)
            if d is None:
                continue
            # patch shared through L2: rr rows x (co-running tiles / row-tiles-per-group) column tiles
            tile_rows = 128 * cg
            groups = 148 // cg
            rows_tiles = max(1, min(rr, size) // tile_rows)
            t_n = rows_tiles * tile_rows
            t_m = max(256.0, groups / rows_tiles * 256.0)
            model = eb * size * size * (1 + size / t_n + size / min(t_m, size))
            r = d["roofline"]
            dram = t.get("dram__bytes_read.sum", 0) + t.get("dram__bytes_write.sum", 0)
            w.writerow(["%s_cg%d_s%d_r%d" % (args.workload, cg, st, rr), "%.6f" % (1e-3 * r["kernel_ms"]),
                        "%.1f" % (1e3 * r["achieved"]), cg, st, rr, int(dram), int(model),
                        "%.1f" % t.get("lts__t_sector_hit_rate.pct", float("nan")), d["clocks"]["sm_mhz"]])
            f.flush()
            print(cg, st, rr, "%.3f ms" % r["kernel_ms"], "%.0f GOp/s" % (1e3 * r["achieved"]),
                  "dram %.1f GB (model %.1f GB)" % (dram * 1e-9, model * 1e-9), flush=True)


if __name__ == "__main__":
    main()
