#!/bin/bash
# r01 refresh after the TMA-fed DMMA kernel: full GPU suite, double benches, ncu capture, tile-height
# sweep on the row blocks of a G-GPU split, sanitizer passes over both tile heights
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f achieved %.2f frac %.3f | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["achieved"], r["frac"], c["sm_mhz"], c["power_w_max"], c["reasons"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --workload double8192 > $O/bench_double8192_default.json 2>$O/bench_double8192_default.err; tail -1 $O/bench_double8192_default.json | python -c "$J" "double8192 default"
timeout 600 python bench.py --workload double8192 --steps 30 --no-e2e --no-cpu > $O/bench_double8192_sustained.json 2>/dev/null; tail -1 $O/bench_double8192_sustained.json | python -c "$J" "double8192 x30"
echo "== tile heights on row blocks"; timeout 600 python scripts/exp_dmma_rows.py 2>&1 | tee gpurun_out/exp_dmma_rows2.log
echo "== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -f -o $O/ncu_dmma python bench.py --workload double8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
for tool in memcheck racecheck; do
  for rows in 64 128; do
    echo "== compute-sanitizer $tool, MM_DMMA_TILE_ROWS=$rows"
    SANITIZE_ONLY=dmma MM_DMMA_TILE_ROWS=$rows timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_dmma_${tool}_$rows.log 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$|MISMATCH|Error|hazard" $O/sanitizer_dmma_${tool}_$rows.log | sort | uniq -c | head
  done
done
