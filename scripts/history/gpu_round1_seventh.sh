#!/bin/bash
# Final r01 refresh: suite, default + sustained benches for all four workloads with the final kernels,
# CPU baseline of the reference path on this host, host executables (timing after the dry-run fix)
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.1f frac %.3f | e2e %s | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], (d.get("e2e") or {}).get("value"), c["sm_mhz"], c["power_w_max"], c["reasons"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in float16384 half32768 double8192 addmin8192; do
  timeout 900 python bench.py --workload $wl > $O/bench_${wl}_default.json 2>$O/bench_${wl}_default.err; tail -1 $O/bench_${wl}_default.json | python -c "$J" "$wl default"
  steps=100; [ $wl != float16384 ] && steps=30
  timeout 900 python bench.py --workload $wl --steps $steps --no-e2e --no-cpu > $O/bench_${wl}_sustained.json 2>/dev/null; tail -1 $O/bench_${wl}_sustained.json | python -c "$J" "$wl x$steps"
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 > $O/bench_reference_arm.json 2>/dev/null; tail -1 $O/bench_reference_arm.json | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_half32768.csv python bench.py --workload half32768 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_f16 python bench.py --workload half32768 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "ncu f16 rc=$?"
echo "== cpu baseline"; timeout 900 python scripts/cpu_baseline.py 2>/dev/null | tail -1 > $O/cpu_baseline.json; python -c "
import json; d=json.load(open('$O/cpu_baseline.json')); print(d['cpu_model'], d['host_cpus'])
for r in d['naive']: print(r['config'], r['shape'], '%.3f GOp/s'%r['gops'])
for r in d['test_simulation']: print('TestSimulation', r['config'], '%.2f s'%r['seconds'], r['verified'])"
echo "== host executables"
mkdir -p /tmp/hostbuild && cd /tmp/hostbuild && python - <<'PY'
import re,os
root=os.environ.get("GRAFT_REPO_ROOT","/root/repo")
t=open(root+"/gemm_hls_b200/host/Config.h.in").read()
cfg=dict(MM_HOST_DATA_TYPE="float",MM_DATA_TYPE="float",MM_DTYPE_CODE="MM_DTYPE_FLOAT",MM_MAP_OP_UPPER="MULTIPLY",MM_MAP_OP="Multiply",MM_REDUCE_OP_UPPER="ADD",MM_REDUCE_OP="Add",MM_MEMORY_BUS_WIDTH_K=64,MM_MEMORY_BUS_WIDTH_M=64,MM_SIZE_N=512,MM_SIZE_K=512,MM_SIZE_M=512,MM_MEMORY_TILE_SIZE_N=128,MM_MEMORY_TILE_SIZE_M=256)
t=re.sub(r"\$\{(\w+)\}",lambda m:str(cfg[m.group(1)]),t).replace("#cmakedefine MM_EXACT","/* #undef MM_EXACT */")
open("Config.h","w").write(t)
PY
R=${GRAFT_REPO_ROOT:-/root/repo}
for exe in TestSimulation RunHardware PrintSpecifications; do
  src=$R/gemm_hls_b200/host/$exe.cpp; extra=""; [ $exe = TestSimulation ] && extra=$R/gemm_hls_b200/host/KernelEntry.cpp
  g++ -std=c++17 -O2 -DMM_DYNAMIC_SIZES -I. -I$R/include -I$R/gemm_hls_b200/host $src $extra -L$R/gemm_hls_b200 -lmm_b200 -Wl,-rpath,$R/gemm_hls_b200 -ldl -lpthread -o $exe || echo "build of $exe failed"
done
cd $R
( /tmp/hostbuild/TestSimulation 513 528 528; echo "TestSimulation rc=$?"; /tmp/hostbuild/RunHardware 1024 1024 1024 hw on; echo "RunHardware rc=$?"; /tmp/hostbuild/RunHardware 16384 16384 16384 hw off; echo "RunHardware rc=$?"; /tmp/hostbuild/PrintSpecifications 16384 16384 16384 ) > $O/host_executables.log 2>&1; grep -E "rc=|Kernel executed|verified" $O/host_executables.log
