#!/bin/bash
# Final r01 refresh after the fused A rounding: suite, float benches, launch list, ncu capture, sanitizers
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.1f frac %.3f | e2e %s | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], (d.get("e2e") or {}).get("value"), c["sm_mhz"], c["power_w_max"], c["reasons"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_float16384_default.json 2>$O/bench_float16384_default.err; tail -1 $O/bench_float16384_default.json | python -c "$J" "float16384 default"
timeout 900 python bench.py --steps 100 --no-e2e --no-cpu > $O/bench_float16384_sustained.json 2>/dev/null; tail -1 $O/bench_float16384_sustained.json | python -c "$J" "float16384 x100"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_tf32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "ncu rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY" $O/sanitizer_$tool.log
done
