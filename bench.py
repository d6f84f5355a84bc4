#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json: "GFLOP/s at N=K=M=16384 fp32").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One step = one MatrixMultiplicationKernel invocation C = A * B (operand preparation + GEMM, or the
configured semiring) over one batch of synthetic matrices through the C-ABI of libmm_b200.so.
For N > 1 (torchrun, one rank per GPU, NCCL) the C row-blocks are split across ranks, B is
broadcast ONCE from rank 0 before the timed region (no per-step collective), every
rank multiplies its row-block each step; time = max over ranks, value = total ops / time.

Printed JSON line (rank 0): see the contract in the task statement; in addition
  roofline      dominant kernel's achieved rate vs the measured peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own Naive<> (oracle/_ref, include/Utility.h:18-42) timed on this
                host's cores on a bounded sample of the same workload (rank 0, N == 1)
  e2e           the same metric through mm_gemm_host() with HOST (pinned) buffers, H2D + D2H inside
`--impl reference` times only the reference CPU path (oracle/_ref; the oracle port if absent).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dtype name, map, reduce, n, k, m, BASELINE.json config it is)
    "float16384": ("float", "Multiply", "Add", 16384, 16384, 16384, "configs[1] float 16384^3 tcgen05"),
    "half32768": ("half", "Multiply", "Add", 32768, 32768, 32768, "configs[2] half 32768^3"),
    "double8192": ("double", "Multiply", "Add", 8192, 8192, 8192, "configs[3] double 8192^3"),
    "addmin8192": ("float", "Add", "Min", 8192, 8192, 8192, "configs[4] (add,min) float 8192^3"),
    "float4096": ("float", "Multiply", "Add", 4096, 4096, 4096, "reduced size, debugging only"),
}
DEFAULT_WORKLOAD = "float16384"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        p["_source"] = "measured"
        return p
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # keep the samples taken under load (power above the idle floor) when there are any
        loaded = [s for s, p in zip(sm, power) if p > 300.0] or sm
        loaded_power = [p for p in power if p > 300.0]
        return {"sm_mhz": statistics.median(loaded) if loaded else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None,
                "power_w_avg_under_load": (sum(loaded_power) / len(loaded_power)) if loaded_power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """Host threads for the reference's CPU path: one per PHYSICAL core this process may run on.

    Naive<> walks a column of B with a stride of M elements: 16384 cache lines (1 MiB) per output element,
    reused by the next 15 columns.  That working set fits one core's private L2 once, not twice, so two
    hyper-threads on a core evict each other (on the 128-thread host of the B200 boxes a float 16384^2 step
    with one row on each of the 128 logical CPUs did not finish within 45 s; one row on one thread takes 5 s)."""
    try:
        allowed = set(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        allowed = None
    cores, cpu, pkg = set(), None, None
    try:
        for line in open("/proc/cpuinfo"):
            key, _, val = line.partition(":")
            key, val = key.strip(), val.strip()
            if key == "processor":
                cpu, pkg = int(val), None
            elif key == "physical id":
                pkg = val
            elif key == "core id" and cpu is not None and (allowed is None or cpu in allowed):
                cores.add((pkg, val))
    except OSError:
        pass
    if cores:
        return len(cores)
    return max(1, len(allowed) if allowed else (os.cpu_count() or 1))


def reference_naive_sample(dtype_name, mp_name, rd_name, a_rows, b, k, m, threads=1):
    """Time the reference's own Naive<> (oracle/_ref) on the rows of C that `a_rows` selects.

    Naive<> is single-threaded as written (include/Utility.h:18-42); C rows are independent, so
    `threads` host threads each run the reference's unmodified routine on their own share of the rows
    (ctypes releases the GIL during the call).  Returns (wall seconds, kind, threads used)."""
    from concurrent.futures import ThreadPoolExecutor
    import oracle as O
    dt = {"float": O.FLOAT, "half": O.HALF, "double": O.DOUBLE}[dtype_name]
    mp, rd = getattr(O, mp_name.upper()), getattr(O, rd_name.upper())
    rows = a_rows.shape[0]
    threads = max(1, min(threads, rows))
    bounds = [rows * i // threads for i in range(threads + 1)]
    use_ref = O.ref_available(dt, mp, rd)
    if use_ref:
        O.ref_lib(dt, mp, rd)   # load once, before the threads race for it
    else:
        O.lib()

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        if use_ref:
            O.ref_naive(dt, mp, rd, a_rows[lo:hi], b, hi - lo, k, m)
        else:
            O.naive(dt, mp, rd, a_rows[lo:hi], b, hi - lo, k, m, threads=1)

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(threads)))
    return time.perf_counter() - t0, ("reference" if use_ref else "port"), threads


def cpu_baseline_line(dtype_name, mp_name, rd_name, unit, k, m, a_rows_of, b):
    """The `cpu_baseline` object of the B200 arm: the reference's Naive<> on a bounded sample of the same
    workload — one row of C per host thread (two when that still stays under ~3e9 operations in total).
    `a_rows_of(rows)` returns the first rows of A as a host array (at most `rows`), `b` is B on the host."""
    threads = host_threads()
    a_rows = a_rows_of(threads * (2 if 2.0 * 2 * threads * k * m <= 3e9 else 1))
    sample_rows = a_rows.shape[0]
    secs, kind, threads = reference_naive_sample(dtype_name, mp_name, rd_name, a_rows, b, k, m, threads)
    return {"value": 1e-9 * 2.0 * sample_rows * k * m / secs, "unit": unit, "cores": threads, "kind": kind,
            "seconds": secs, "host_cpus": os.cpu_count(),
            "sample": "first %d rows of C (%d x %d x %d): the reference's Naive<> (include/Utility.h:18-42, "
                      "single-threaded as written) on %d host threads, each on its own rows"
                      % (sample_rows, sample_rows, k, m, threads)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="MM_FLAG_* bits (debugging)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dtype_name, mp_name, rd_name, N, K, M, cfg_label = WORKLOADS[args.workload]
    ops_total = 2.0 * N * K * M
    metric = "GFLOP/s" if (mp_name, rd_name) == ("Multiply", "Add") else "GOp/s"
    metric_name = "%s at N=%d K=%d M=%d %s (%s,%s)" % (metric, N, K, M, dtype_name, mp_name, rd_name)
    config = {"workload": "%s %dx%dx%d (%s,%s)" % (dtype_name, N, K, M, mp_name, rd_name), "baseline_config": cfg_label,
              "partition": "C row-blocks over %d GPU(s), B replicated (one NCCL broadcast before timing)" % world,
              "l2": "inputs (A+B+C = %.2f GB) far larger than the 126 MB L2; no explicit flush" %
                    (1e-9 * {"float": 4, "half": 2, "double": 8}[dtype_name] * (N * K + K * M + N * M))}

    import numpy as np
    np_dt = {"float": np.float32, "half": np.float16, "double": np.float64}[dtype_name]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # one step = one row of C per host thread (C rows are independent; every thread runs the
        # reference's single-threaded Naive<> on its own row): ~5 s of wall clock at 16384^2 per row
        threads = host_threads()
        rows_per_step = threads
        rng = np.random.default_rng(5)
        b = rng.uniform(1, 10, size=(K, M)).astype(np_dt)
        a_rows = rng.uniform(1, 10, size=(rows_per_step, K)).astype(np_dt)
        kind = "reference"
        # one untimed step; if the host is slower than expected, shrink the per-step sample (fewer rows,
        # fewer threads) so that K timed steps stay within a few minutes
        t_warm, _, _ = reference_naive_sample(dtype_name, mp_name, rd_name, a_rows, b, K, M, threads)
        if t_warm > 8.0:
            threads = max(1, int(threads * 8.0 / t_warm))
            rows_per_step = threads
            a_rows = a_rows[:rows_per_step]
        t = 0.0
        for _ in range(args.steps):
            dt_s, kind, threads = reference_naive_sample(dtype_name, mp_name, rd_name, a_rows, b, K, M, threads)
            t += dt_s
        sample_ops = 2.0 * rows_per_step * K * M
        value = 1e-9 * sample_ops * args.steps / t
        sample = ("%d rows of C per step (%d x %d x %d): the reference's Naive<> (single-threaded as written) on one "
                  "row per host thread, %d threads" % (rows_per_step, rows_per_step, K, M, threads))
        print(json.dumps({
            "impl": "reference", "metric": metric_name, "value": value, "unit": metric, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype_name,
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": metric, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": metric, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import gemm_hls_b200 as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    t_dt = {"float": torch.float32, "half": torch.float16, "double": torch.float64}[dtype_name]
    dtype = G.DTYPE_FROM_NAME[dtype_name]
    mp, rd = G.OP_FROM_NAME[mp_name], G.OP_FROM_NAME[rd_name]
    es = torch.empty((), dtype=t_dt).element_size()

    # row-block of this rank: ceil(N / world) rows, last block may be short
    rows_per = (N + world - 1) // world
    r0, r1 = min(N, rank * rows_per), min(N, (rank + 1) * rows_per)
    n_local = r1 - r0

    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)
    # synthetic U[1,10) inputs as in the reference recipe (test/TestSimulation.cpp:46-55); half uses
    # U[0,1) so that C stays finite in half (SURVEY.md trap 5)
    lo, hi = (0.0, 1.0) if dtype_name == "half" else (1.0, 10.0)
    a_blk = (torch.rand((n_local, K), generator=gen, device=dev, dtype=torch.float32) * (hi - lo) + lo).to(t_dt)
    if rank == 0:
        b_full = (torch.rand((K, M), generator=gen, device=dev, dtype=torch.float32) * (hi - lo) + lo).to(t_dt)
    else:
        b_full = torch.empty((K, M), device=dev, dtype=t_dt)
    if world > 1:
        dist.broadcast(b_full, src=0)  # the ONE collective of the path: B over NVLink/NVSwitch
        # reported beside the step time: the same broadcast once more, now that the
        # communicator exists, timed on the device, max over ranks
        torch.cuda.synchronize()
        dist.barrier()
        eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eb0.record()
        dist.broadcast(b_full, src=0)
        eb1.record()
        torch.cuda.synchronize()
        tb = torch.tensor([eb0.elapsed_time(eb1)], device=dev, dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        config["broadcast_b_ms"] = round(float(tb.item()), 4)
        config["broadcast_b_bytes"] = b_full.numel() * b_full.element_size()
    c_blk = torch.empty((n_local, M), device=dev, dtype=t_dt)
    torch.cuda.synchronize()

    ctx = G.Context(local_rank)
    # a dedicated (non-default) torch stream: its handle is what the C-ABI launches on and what the
    # torch.cuda.Event pairs below are recorded on (the default stream's handle is 0 == "use the
    # context's own stream" in the C-ABI)
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream
    assert stream != 0
    flags = args.flags

    def step():
        ctx.enqueue(dtype, mp, rd, a_blk.data_ptr(), b_full.data_ptr(), c_blk.data_ptr(), n_local, K, M,
                    flags=flags, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ctx.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    prep_s, main_s, calls = ctx.profile_read()
    ctx.set_profiling(False)
    clocks = sampler.stop() if sampler else None

    t_max = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_max.item())
    ms_per_step = elapsed_ms / args.steps
    value = 1e-9 * ops_total / (1e-3 * ms_per_step)  # whole job: all ranks' row-blocks

    # ---- light on-device sanity so that a wrong kernel cannot post a number (not the parity test)
    if (mp_name, rd_name) == ("Multiply", "Add"):
        rows = torch.tensor([0, n_local // 2, n_local - 1], device=dev)
        ref = a_blk[rows].double() @ b_full.double()
        got = c_blk[rows].double()
        rel = ((got - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()
        tol = 1e-2 if dtype_name == "half" else 1e-3
        if not (rel <= tol):
            raise SystemExit("bench.py: result check failed (max rel err %.3e > %.0e)" % (rel, tol))
        config["check"] = "3 rows of C vs fp64 torch.matmul on device: max rel err %.2e" % rel

    out = None
    if rank == 0:
        peaks = load_peaks()
        path = G.kernel_path(dtype, mp, rd, flags)
        main_avg_s = main_s / max(calls, 1)
        local_ops = 2.0 * n_local * K * M
        if path in ("tcgen05_tf32", "tcgen05_f16"):
            # burst figure when the whole timed region is shorter than the ~1 s it takes the power
            # cap to pull the clocks down, the sustained one for a seconds-long back-to-back loop
            long_run = elapsed_ms > 1500.0
            peak_bf16 = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) if long_run else peaks["bf16_tflops"]
            peak = peak_bf16 / 2.0 if path == "tcgen05_tf32" else peak_bf16
            peak_note = ("%s bf16 %s %.1f TF/s%s" % (peaks["_source"], "sustained" if long_run else "burst", peak_bf16,
                         " / 2 (kind::tf32 issues at half the 16-bit rate)" if path == "tcgen05_tf32" else ""))
            roof = {"bound": "tensor", "achieved": 1e-12 * local_ops / main_avg_s, "peak": peak, "unit": "TFLOP/s"}
        elif path == "dmma_f64":
            # FP64 DMMA is not in MEASURED_PEAKS.json.  Measured on this pool with a registers-only DMMA loop
            # (scripts/exp_fp64_pipes.cu, profiles/r01_exp_fp64_pipes.jsonl): 37.05-37.13 TF/s at 1965 MHz
            # = 64 FMA/clk/SM; the HGX B200 datasheet's 296 TF / 8 GPUs = 37 TF/s.
            peak = 37.1
            peak_note = ("FP64 tensor (DMMA) 37.1 TF/s: registers-only DMMA loop measured on this pool "
                         "(profiles/r01_exp_fp64_pipes.jsonl); datasheet 37; not in MEASURED_PEAKS.json")
            roof = {"bound": "tensor", "achieved": 1e-12 * local_ops / main_avg_s, "peak": peak, "unit": "TFLOP/s"}
        else:
            # derived CUDA-core issue ceiling: one warp instruction per clock and scheduler
            # = 148 SMs x 4 x 32 lanes x 1.965 GHz = 37.2e12 lane-instructions/s, 2 ops per element-step.
            #   float (Add, Min|Max): 1 FADD2 + 1 FMNMX3 per two element-steps = 1.0 slot per step -> 74.4 TOp/s
            #     (the half-rate ALU pipe of the FMNMX3 gives the same bound); the inner loop alone, on a
            #     resident tile, measures 44.8 TOp/s (scripts/exp_semiring_issue.cu)
            #   anything else (e.g. float (Multiply, Add) under MM_FLAG_EXACT: 1 FMUL2 per two steps + 1 FADD
            #     per step): 1.5 slots per step -> 49.6
            fast_minmax = dtype_name == "float" and mp_name == "Add" and rd_name in ("Min", "Max") and not (flags & 2)
            peak = 74.4 if fast_minmax else 49.6
            peak_note = ("derived CUDA-core issue ceiling at 1965 MHz, %s (DESIGN.md 3.3)%s; neither HBM- nor "
                         "tensor-bound" % ("1 FADD2 + 1 FMNMX3 per two element-steps" if fast_minmax else
                                           "1.5 issue slots per element-step",
                                           "; inner loop alone measured at 44.8 TOp/s (profiles/r01_exp_semiring_issue.jsonl)"
                                           if fast_minmax else ""))
            roof = {"bound": "cuda_core_issue", "achieved": 1e-12 * local_ops / main_avg_s, "peak": peak, "unit": "TOp/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel"] = path
        roof["kernel_ms"] = 1e3 * main_avg_s
        roof["prep_ms"] = 1e3 * prep_s / max(calls, 1)
        roof["peak_source"] = peak_note
        tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        traffic = None
        if os.path.exists(tr_path):
            traffic = json.load(open(tr_path)).get("%s@%s" % (path, args.workload))
        roof["traffic"] = traffic
        roof["algorithmic_bytes"] = es * (n_local * K + K * M + n_local * M)

        out = {"metric": metric_name, "value": value, "unit": metric, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": {"float": "tf32 tensor-core multiply, f32 accumulate/storage",
                                              "half": "f16 multiply, f32 accumulate", "double": "f64"}[dtype_name]
               if (mp_name, rd_name) == ("Multiply", "Add") else "f32",
               "data": "synthetic", "config": config, "clocks": clocks, "roofline": roof,
               # NVML board power during the timed region (the reference's PSU power meter);
               # meaningful for runs of a second or more (nvidia-smi refreshes every ~100 ms)
               "energy": ({"avg_power_w": clocks["power_w_avg_under_load"],
                           "gop_per_joule": value / clocks["power_w_avg_under_load"]}
                          if clocks and clocks.get("power_w_avg_under_load") else None),
               "gpu_launches": args.steps * G.launch_count(dtype, mp, rd, flags)}

    # ------------------------------------------------------------------ e2e: host buffers through the C-ABI
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        a_host = torch.empty((n_local, K), dtype=t_dt, pin_memory=True)
        b_host = torch.empty((K, M), dtype=t_dt, pin_memory=True)
        c_host = torch.empty((n_local, M), dtype=t_dt, pin_memory=True)
        a_host.copy_(a_blk)
        b_host.copy_(b_full)
        torch.cuda.synchronize()
        a_np, b_np, c_np = a_host.numpy(), b_host.numpy(), c_host.numpy()
        ctx.gemm_host(dtype, mp, rd, a_np, b_np, n_local, K, M, flags=flags, out=c_np)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.gemm_host(dtype, mp, rd, a_np, b_np, n_local, K, M, flags=flags, out=c_np)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        if out is not None:
            out["e2e"] = {"value": 1e-9 * ops_total / float(t_e.item()), "unit": metric,
                          "h2d_bytes_per_step": int(es * (n_local * K + K * M)),
                          "d2h_bytes_per_step": int(es * n_local * M), "steps": e2e_steps,
                          "note": "mm_gemm_host(): pinned host A,B -> device, kernels, C -> pinned host; wall clock, max over ranks"}
        del a_host, c_host

    # ------------------------------------------------------------------ cpu_baseline (rank 0, N == 1)
    if out is not None and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_line(dtype_name, mp_name, rd_name, metric, K, M,
                                                lambda rows: a_blk[:min(rows, n_local)].cpu().numpy(),
                                                b_full.cpu().numpy())

    if out is not None:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
