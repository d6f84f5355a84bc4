#!/bin/bash
# Final-kernel evidence run: tests, benches (burst + sustained), launch list, ncu full captures
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.1f frac %.3f | e2e %s | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], (d.get("e2e") or {}).get("value"), c["sm_mhz"], c["power_w_max"], c["reasons"]))'
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/gpu.txt
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
for wl in float16384 half32768 double8192 addmin8192; do
  timeout 900 python bench.py --workload $wl > $O/bench_${wl}_default.json 2>$O/bench_${wl}_default.err; tail -1 $O/bench_${wl}_default.json | python -c "$J" "$wl default"
  steps=100; [ $wl = half32768 ] && steps=30; [ $wl = double8192 ] && steps=30; [ $wl = addmin8192 ] && steps=30
  timeout 900 python bench.py --workload $wl --steps $steps --no-e2e --no-cpu > $O/bench_${wl}_sustained.json 2>/dev/null; tail -1 $O/bench_${wl}_sustained.json | python -c "$J" "$wl x$steps"
done
timeout 600 python bench.py --flags 4 --steps 10 --no-e2e --no-cpu > $O/bench_float16384_tf32x3.json 2>/dev/null; tail -1 $O/bench_float16384_tf32x3.json | python -c "$J" "float16384 tf32x3"
timeout 600 python bench.py --flags 2 --steps 3 --no-e2e --no-cpu > $O/bench_float16384_exact.json 2>/dev/null; tail -1 $O/bench_float16384_exact.json | python -c "$J" "float16384 exact(simt)"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 > $O/bench_reference_arm.json 2>/dev/null; tail -1 $O/bench_reference_arm.json | cut -c1-300
echo "== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_half32768.csv python bench.py --workload half32768 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
echo "== ncu full captures"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_tf32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o $O/ncu_tcgen05_f16 python bench.py --workload half32768 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"transpose_prep|round_tf32" -s 2 -c 2 -f -o $O/ncu_prep python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o $O/ncu_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -f -o $O/ncu_dmma python bench.py --workload double8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== host executables on the GPU (cmake build is not shipped; build the three with g++ against the in-tree .so)"
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
( /tmp/hostbuild/TestSimulation 513 528 528; echo "TestSimulation rc=$?"; /tmp/hostbuild/RunHardware 2048 2048 2048 hw on; echo "RunHardware rc=$?"; /tmp/hostbuild/RunHardware 16384 16384 16384 hw off; echo "RunHardware rc=$?"; /tmp/hostbuild/PrintSpecifications 16384 16384 16384 ) > $O/host_executables.log 2>&1; tail -22 $O/host_executables.log
