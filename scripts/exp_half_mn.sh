#!/bin/bash
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_half_mn.log; : > $out
echo "== full gpu suite (half reads B MN-major)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee -a $out
echo "== single-CTA variant, half subset"; MM_TCGEN05_CTA_GROUP=1 timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "half or golden or transposed" 2>&1 | tail -2 | tee -a $out
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-26s ms/step %.3f kernel_ms %.3f prep_ms %.3f TF(step) %.1f TF(kernel) %.1f | launches %d | sm_mhz %s power %s" % (sys.argv[1], d["ms_per_step"], r["kernel_ms"], r["prep_ms"], d["value"]*1e-3, r["achieved"], d["gpu_launches"], c["sm_mhz"], c["power_w_max"]))'
run() { label=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "$J" "$label" | tee -a $out; }
run "f16 B MN-major x30" half32768 30 MM_TCGEN05_B_MN=1
run "f16 B transposed x30" half32768 30 MM_TCGEN05_B_MN=0
run "f16 B MN-major x10" half32768 10 MM_TCGEN05_B_MN=1
run "f16 B transposed x10" half32768 10 MM_TCGEN05_B_MN=0
echo "== cpu baseline of the reference path on this host"; timeout 900 python scripts/cpu_baseline.py 2>/dev/null | tail -1 > gpurun_out/r01/cpu_baseline.json; python -c "
import json; d=json.load(open('gpurun_out/r01/cpu_baseline.json')); print(d['cpu_model'], d['host_cpus'])
for r in d['naive']: print(r['config'], r['shape'], '%.3f GOp/s'%r['gops'])
for r in d['test_simulation']: print('TestSimulation', r['config'], '%.2f s'%r['seconds'], r['verified'])" | tee -a $out
