#!/bin/bash
# r01 refresh after the packed FADD2 / FMUL2 path of the semiring kernel: full GPU suite, (add,min) and
# exact-float benches, ncu capture of the (add,min) kernel, sanitizer pass over the float semiring cases
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f achieved %.2f frac %.3f | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["achieved"], r["frac"], c["sm_mhz"], c["power_w_max"], c["reasons"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --workload addmin8192 > $O/bench_addmin8192_default.json 2>$O/bench_addmin8192_default.err; tail -1 $O/bench_addmin8192_default.json | python -c "$J" "addmin8192 default"
timeout 600 python bench.py --workload addmin8192 --steps 30 --no-e2e --no-cpu > $O/bench_addmin8192_sustained.json 2>/dev/null; tail -1 $O/bench_addmin8192_sustained.json | python -c "$J" "addmin8192 x30"
timeout 600 python bench.py --flags 2 --steps 3 --no-e2e --no-cpu > $O/bench_float16384_exact.json 2>/dev/null; tail -1 $O/bench_float16384_exact.json | python -c "$J" "float16384 exact(simt)"
echo "== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o $O/ncu_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== compute-sanitizer memcheck (float semiring cases)"
SANITIZE_ONLY="semiring f32" timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_semiring_f32_memcheck.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|ok$|MISMATCH|Error" $O/sanitizer_semiring_f32_memcheck.log | sort | uniq -c | head
