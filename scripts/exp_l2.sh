#!/bin/bash
# Diagnostics: DRAM traffic / duration of the tcgen05 kernel under ring-depth and L2-policy variants
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_l2.log; : > $out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv python bench.py --workload ${WL:-float16384} --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "dram__bytes_read|gpu__time|hit_rate|pipe_tensor" | awk -F'","' '{printf "%s=%s%s  ", $(NF-2), $NF, $(NF-1)}' | sed 's/"//g' | sed "s/^/$label: /" | tee -a $out; echo | tee -a $out
}
run "cg1 s4 normal" MM_TCGEN05_CTA_GROUP=1
run "cg1 s4 last  " MM_TCGEN05_CTA_GROUP=1 MM_TCGEN05_L2=last
run "cg1 s3 normal" MM_TCGEN05_CTA_GROUP=1 MM_TCGEN05_STAGES=3
run "cg2 s6 normal" MM_TCGEN05_CTA_GROUP=2
run "cg2 s6 last  " MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_L2=last
run "cg2 s4 normal" MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4
run "cg2 s3 normal" MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=3
run "cg2 s4 last  " MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4 MM_TCGEN05_L2=last
WL=half32768 run "f16 cg1 s4 normal" MM_TCGEN05_CTA_GROUP=1
WL=half32768 run "f16 cg2 s6 normal" MM_TCGEN05_CTA_GROUP=2
WL=half32768 run "f16 cg2 s4 last  " MM_TCGEN05_CTA_GROUP=2 MM_TCGEN05_STAGES=4 MM_TCGEN05_L2=last
echo "== tf32x3 + quick tests"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "tf32x3 or float_tensor or golden" 2>&1 | tail -3 | tee -a $out
