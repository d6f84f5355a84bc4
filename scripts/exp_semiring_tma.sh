#!/bin/bash
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f achieved %.2f frac %.3f | sm_mhz %s power %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["achieved"], r["frac"], c["sm_mhz"], c["power_w_max"]))'
for i in 1 2; do timeout 600 python bench.py --workload addmin8192 --steps 10 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "$J" "addmin8192 (TMA B)"; done
timeout 600 python bench.py --workload addmin8192 > $O/bench_addmin8192_default.json 2>/dev/null; tail -1 $O/bench_addmin8192_default.json | python -c "$J" "addmin8192 default"
timeout 600 python bench.py --flags 2 --steps 3 --no-e2e --no-cpu > $O/bench_float16384_exact.json 2>/dev/null; tail -1 $O/bench_float16384_exact.json | python -c "$J" "float16384 exact(simt)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o $O/ncu_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "ncu rc=$?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY" $O/sanitizer_$tool.log
done
