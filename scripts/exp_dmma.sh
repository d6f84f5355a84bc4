#!/bin/bash
# Ran against the LDGSTS-fed DMMA kernel (before commit ad15438); MM_DMMA_WARPS no longer exists.
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_dmma.log; : > $out
echo "== double tests (16 warps default)"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "double or DOUBLE or golden" 2>&1 | tail -2 | tee -a $out
echo "== double tests (8 warps)"; MM_DMMA_WARPS=8 timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "double_tensor" 2>&1 | tail -2 | tee -a $out
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-22s ms/step %.3f TF %.2f frac %.3f | sm_mhz %s power %s" % (sys.argv[1], d["ms_per_step"], r["achieved"], r["frac"], c["sm_mhz"], c["power_w_max"]))'
for w in 16 8 16 8; do
  MM_DMMA_WARPS=$w timeout 600 python bench.py --workload double8192 --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "$J" "double8192 warps=$w" | tee -a $out
done
