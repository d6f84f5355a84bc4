#!/bin/bash
# One-stop evidence run on a B200 box (under gpurun):  bash scripts/gpu_evidence.sh [TAG] [section ...]
#   TAG       output goes to gpurun_out/TAG/ (default r01); afterwards, here: python scripts/make_profiles.py TAG
#   sections  suite smoke bench sustained variants reference launches ncu cpu host sanitize  (default: all)
# Every stage runs in its own process under a timeout, so a hung kernel cannot take the whole call down.
# Rough box time for everything on one GPU: 12-15 minutes (the ncu captures are about half of it).
set +e
TAG=${1:-r01}; shift
SECTIONS=${*:-suite smoke bench sustained variants reference launches ncu cpu host sanitize}
R=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd "$R"
O=gpurun_out/$TAG
mkdir -p $O
WORKLOADS="float16384 half32768 double8192 addmin8192"
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; e=d.get("e2e") or {}; print("%-24s ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.2f frac %.3f | e2e %s | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], e.get("value"), c["sm_mhz"], c["power_w_max"], c["reasons"]))'
has() { [[ " $SECTIONS " == *" $1 "* ]]; }

nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/gpu.txt

if has suite; then
  echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
fi
if has smoke; then
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
fi
if has bench; then
  for wl in $WORKLOADS; do
    timeout 900 python bench.py --workload $wl > $O/bench_${wl}_default.json 2>$O/bench_${wl}_default.err
    tail -1 $O/bench_${wl}_default.json | python -c "$J" "$wl default"
  done
fi
if has sustained; then
  for wl in $WORKLOADS; do
    steps=30; [ $wl = float16384 ] && steps=100
    timeout 900 python bench.py --workload $wl --steps $steps --no-e2e --no-cpu > $O/bench_${wl}_sustained.json 2>/dev/null
    tail -1 $O/bench_${wl}_sustained.json | python -c "$J" "$wl x$steps"
  done
fi
if has variants; then
  timeout 600 python bench.py --flags 4 --steps 10 --no-e2e --no-cpu > $O/bench_float16384_tf32x3.json 2>/dev/null; tail -1 $O/bench_float16384_tf32x3.json | python -c "$J" "float16384 tf32x3"
  timeout 600 python bench.py --flags 2 --steps 3 --no-e2e --no-cpu > $O/bench_float16384_exact.json 2>/dev/null; tail -1 $O/bench_float16384_exact.json | python -c "$J" "float16384 exact(simt)"
  timeout 600 python scripts/library_baselines.py > $O/library_baselines.json 2>/dev/null; head -c 600 $O/library_baselines.json
fi
if has reference; then
  echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 > $O/bench_reference_arm.json 2>/dev/null; tail -1 $O/bench_reference_arm.json | cut -c1-300
fi
if has launches; then
  echo "== ncu launch lists (gpu__time_duration per launch of the bench command)"
  for wl in float16384 half32768; do
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  done
fi
if has ncu; then
  echo "== ncu --set full, one launch of each kernel"
  cap() {  # name, kernel regex, skip, count, bench args...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o $O/$name python bench.py "$@" --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "$name rc=$?"
  }
  cap ncu_tcgen05_tf32 gemm_tcgen05 1 1
  cap ncu_tcgen05_f16 gemm_tcgen05 1 1 --workload half32768
  cap ncu_prep "transpose_prep|round_tf32" 2 2
  cap ncu_semiring_addmin semiring_tile 1 1 --workload addmin8192
  cap ncu_dmma gemm_dmma 1 1 --workload double8192
fi
if has cpu; then
  echo "== reference CPU path on this host"; timeout 900 python scripts/cpu_baseline.py 2>/dev/null | tail -1 > $O/cpu_baseline.json; head -c 400 $O/cpu_baseline.json; echo
fi
if has host; then
  echo "== host executables (scripts/build_host.sh, single-GPU programs)"
  bash scripts/build_host.sh /tmp/hostbuild > /dev/null 2>&1 || echo "host build failed"
  ( /tmp/hostbuild/TestSimulation 513 528 528; echo "TestSimulation rc=$?"
    /tmp/hostbuild/RunHardware 1024 1024 1024 hw on; echo "RunHardware rc=$?"
    MM_POWER_METER=1 /tmp/hostbuild/RunHardware 16384 16384 16384 hw off; echo "RunHardware(power meter) rc=$?"
    /tmp/hostbuild/PrintSpecifications 16384 16384 16384; echo "PrintSpecifications rc=$?" ) > $O/host_executables.log 2>&1
  grep -E "rc=|Kernel executed|verified|Mismatch" $O/host_executables.log
fi
if has sanitize; then
  for tool in memcheck racecheck synccheck; do
    echo "== compute-sanitizer $tool"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MISMATCH|Error|hazard" $O/sanitizer_$tool.log | sort | uniq -c | head -8
  done
  SANITIZE_ONLY=dmma MM_DMMA_TILE_ROWS=128 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_dmma_memcheck_128.log 2>&1
  grep -E "ERROR SUMMARY" $O/sanitizer_dmma_memcheck_128.log
fi
