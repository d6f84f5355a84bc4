#!/usr/bin/env python
"""Turn the scratch outputs of a GPU evidence run (gpurun_out/rNN/) into the tracked summaries under
profiles/ (read here; no GPU needed):

    python scripts/make_profiles.py r01

  profiles/rNN_bench.jsonl              every bench.py JSON line of the run, one per line, labelled
  profiles/rNN_launches_<wl>.csv        ncu launch list (gpu__time_duration per launch) of bench.py
  profiles/rNN_launch_shares.md         per-kernel share of a step from that list vs the live CUDA-event split
  profiles/rNN_ncu_<kernel>.txt         key metrics of the `ncu --set full` capture of each kernel
  profiles/rNN_ncu_<kernel>_raw.csv     its complete raw metric page
  profiles/ncu_traffic.json             dram bytes per launch for bench.py's roofline.traffic
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_summary  # noqa: E402,F401  (KEYS)


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return out, list(csv.reader(out.splitlines()))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)

    # ---- bench lines
    with open(os.path.join(dst, tag + "_bench.jsonl"), "w") as f:
        for p in sorted(glob.glob(os.path.join(src, "bench_*.json"))):
            lines = [l for l in open(p) if l.startswith("{")]
            if lines:
                d = json.loads(lines[-1])
                d["_run"] = os.path.basename(p)[len("bench_"):-len(".json")]
                f.write(json.dumps(d) + "\n")

    # ---- launch lists + shares
    shares = ["# %s — kernel shares of one bench step (ncu launch list, cold-cache serialised launches)\n" % tag]
    for p in sorted(glob.glob(os.path.join(src, "launches_*.csv"))):
        wl = os.path.basename(p)[len("launches_"):-len(".csv")]
        rows = [r for r in csv.reader(l for l in open(p) if l.startswith('"'))]
        hdr, body = rows[0], rows[1:]
        kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
        with open(os.path.join(dst, "%s_launches_%s.csv" % (tag, wl)), "w") as f:
            w = csv.writer(f)
            w.writerow(["launch", "kernel", "gpu__time_duration_ns"])
            for i, r in enumerate(body):
                w.writerow([i, r[kn][:120], r[val]])
        mine = {}
        for r in body:
            name = r[kn]
            for key in ("gemm_tcgen05_kernel", "transpose_prep_kernel", "round_tf32_kernel", "gemm_dmma",
                        "semiring_tile_kernel", "split3"):
                if key in name:
                    mine.setdefault(key, []).append(float(r[val].replace(",", "")))
        tot = sum(sum(v) / len(v) for v in mine.values()) or 1.0
        shares.append("\n## %s\n\n| kernel | launches | mean ns | share of step |\n|---|---|---|---|" % wl)
        for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
            shares.append("| %s | %d | %.0f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100.0 * (sum(v) / len(v)) / tot))
        bj = os.path.join(src, "bench_%s_default.json" % wl)
        if os.path.exists(bj):
            d = json.loads([l for l in open(bj) if l.startswith("{")][-1])
            r = d["roofline"]
            live = r["kernel_ms"] / (r["kernel_ms"] + r["prep_ms"])
            shares.append("\nlive CUDA-event split of the same step in bench.py: main kernel %.3f ms, preparation %.3f ms "
                          "-> main-kernel share %.1f %%" % (r["kernel_ms"], r["prep_ms"], 100 * live))
    open(os.path.join(dst, tag + "_launch_shares.md"), "w").write("\n".join(shares) + "\n")

    # ---- ncu full captures
    traffic = {}
    tpath = os.path.join(dst, "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    wl_of = {"ncu_tcgen05_tf32": ("tcgen05_tf32", "float16384"), "ncu_tcgen05_f16": ("tcgen05_f16", "half32768"),
             "ncu_dmma": ("dmma_f64", "double8192"), "ncu_semiring_addmin": ("semiring_simt", "addmin8192")}
    for p in sorted(glob.glob(os.path.join(src, "ncu_*.ncu-rep"))):
        base = os.path.basename(p)[:-len(".ncu-rep")]
        text, rows = raw_page(p)
        open(os.path.join(dst, "%s_%s_raw.csv" % (tag, base)), "w").write(text)
        hdr, units = rows[0], rows[1]
        lines = []
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            for k in ncu_summary.KEYS:
                if k in d:
                    lines.append("%-84s %s %s" % (k, d[k][:140], units[hdr.index(k)]))
            lines.append("--")
            if base in wl_of:
                def num(key):
                    v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                path, wl = wl_of[base]
                traffic["%s@%s" % (path, wl)] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
        open(os.path.join(dst, "%s_%s.txt" % (tag, base)), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)

    extras = ["pytest_gpu.log", "host_executables.log", "gpu.txt", "cpu_baseline.json"]
    for pat in ("scale_*.json", "multi_runhardware_*.log", "tile_sweep_*.csv"):
        extras += [os.path.basename(x) for x in glob.glob(os.path.join(src, pat))]
    for extra in extras:
        p = os.path.join(src, extra)
        if os.path.exists(p):
            open(os.path.join(dst, "%s_%s" % (tag, extra)), "w").write(open(p).read())
    # diagnostics written by the scripts/exp_*.sh experiments
    for p in glob.glob(os.path.join(ROOT, "gpurun_out", "exp_*")):
        open(os.path.join(dst, "%s_%s" % (tag, os.path.basename(p))), "w").write(open(p).read())
    print("profiles/ updated from", src)


if __name__ == "__main__":
    main()
