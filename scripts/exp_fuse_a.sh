#!/bin/bash
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_fuse_a3.log; : > $out
echo "== tests"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "float or tf32 or golden or row_block or graph or lifecycle" 2>&1 | tail -2 | tee -a $out
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-24s ms/step %.3f step TF %.1f | kernel_ms %.3f prep_ms %.3f | sm_mhz %s power %s" % (sys.argv[1], d["ms_per_step"], d["value"]*1e-3, r["kernel_ms"], r["prep_ms"], c["sm_mhz"], c["power_w_max"]))'
run() { label=$1; steps=$2; shift 2; env "$@" timeout 600 python bench.py --steps $steps --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "$J" "$label" | tee -a $out; }
for i in 1 2 3; do
run "fused+head x10" 10 MM_TCGEN05_FUSE_A=1
run "separate x10" 10 MM_TCGEN05_FUSE_A=0
done
run "fused+head x100" 100 MM_TCGEN05_FUSE_A=1
run "separate x100" 100 MM_TCGEN05_FUSE_A=0
