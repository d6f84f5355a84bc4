#!/bin/bash
# Evidence refresh for the final tcgen05 configuration (pairs, 6 stages, soft wave barrier) + tile sweep
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.1f frac %.3f | e2e %s | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], (d.get("e2e") or {}).get("value"), c["sm_mhz"], c["power_w_max"], c["reasons"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu.log
for wl in float16384 half32768; do
  timeout 900 python bench.py --workload $wl > $O/bench_${wl}_default.json 2>$O/bench_${wl}_default.err; tail -1 $O/bench_${wl}_default.json | python -c "$J" "$wl default"
  steps=100; [ $wl = half32768 ] && steps=30
  timeout 900 python bench.py --workload $wl --steps $steps --no-e2e --no-cpu > $O/bench_${wl}_sustained.json 2>/dev/null; tail -1 $O/bench_${wl}_sustained.json | python -c "$J" "$wl x$steps"
done
timeout 600 python bench.py --flags 4 --steps 10 --no-e2e --no-cpu > $O/bench_float16384_tf32x3.json 2>/dev/null; tail -1 $O/bench_float16384_tf32x3.json | python -c "$J" "float16384 tf32x3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_half32768.csv python bench.py --workload half32768 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_tf32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_f16 python bench.py --workload half32768 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== tile sweep (half 32768^3, BASELINE config 3)"; timeout 2400 python scripts/tile_sweep.py --workload half32768 --steps 5 --out $O/tile_sweep_half32768.csv 2>&1 | tail -20
