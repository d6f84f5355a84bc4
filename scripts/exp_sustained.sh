#!/bin/bash
# Sustained (power-capped) throughput of the tcgen05 variants: ring depth x CTA group
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_sustained.log; : > $out
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f kernel_ms %.3f TF %.1f | sm_mhz %s power %s %s" % (sys.argv[1], d["ms_per_step"], r["kernel_ms"], r["achieved"], c["sm_mhz"], c["power_w_max"], c["reasons"]))'
run() { label=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "$J" "$label" | tee -a $out; }
echo "== tf32x3 test"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "tf32x3" 2>&1 | grep -E "assert|passed|failed" | head -8 | tee -a $out
for rep in 1 2; do
run "f32 cg1 s4 (rep$rep)" float16384 150 MM_TCGEN05_CTA_GROUP=1
run "f32 cg2 s4 (rep$rep)" float16384 150 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4
run "f32 cg2 s5 (rep$rep)" float16384 150 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=5
run "f32 cg2 s6 (rep$rep)" float16384 150 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=6
done
run "f32 cg1 s4 short" float16384 10 MM_TCGEN05_CTA_GROUP=1
run "f32 cg2 s4 short" float16384 10 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4
run "f32 cg2 s5 short" float16384 10 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=5
run "f32 cg2 s6 short" float16384 10 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=6
run "f16 cg1 s4" half32768 30 MM_TCGEN05_CTA_GROUP=1
run "f16 cg2 s4" half32768 30 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4
run "f16 cg2 s4 last" half32768 30 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4 MM_TCGEN05_L2=last
run "f16 cg2 s5" half32768 30 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=5
run "f16 cg2 s6" half32768 30 MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=6
