#!/bin/bash
# Diagnostics: effect of the soft wave barrier on DRAM re-reads and sustained throughput
set +e
mkdir -p gpurun_out
out=gpurun_out/exp_tile_sync.log; : > $out
echo "== correctness"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "float_tensor or half or tf32 or golden or transposed or row_block" 2>&1 | tail -2 | tee -a $out
ncu_run() {  # label, wl, env...
  label=$1; wl=$2; shift 2
  env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "dram__bytes_read|gpu__time|hit_rate|pipe_tensor" | awk -F'","' '{printf "%s=%s%s  ", $(NF-2), $NF, $(NF-1)}' | sed 's/"//g' | sed "s/^/ncu $label: /" | tee -a $out; echo | tee -a $out
}
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-26s ms/step %.3f kernel_ms %.3f TF %.1f | sm_mhz %s power %s" % (sys.argv[1], d["ms_per_step"], r["kernel_ms"], r["achieved"], c["sm_mhz"], c["power_w_max"]))'
run() { label=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "$J" "$label" | tee -a $out; }
for st in 4 5 6; do
  ncu_run "f32 s$st sync" float16384 MM_TCGEN05_STAGES=$st
  ncu_run "f32 s$st nosync" float16384 MM_TCGEN05_STAGES=$st MM_TCGEN05_TILE_SYNC=0
done
for st in 4 5 6; do
  ncu_run "f16 s$st sync" half32768 MM_TCGEN05_STAGES=$st
done
for st in 4 5 6; do
  run "f32 s$st sync x150" float16384 150 MM_TCGEN05_STAGES=$st
  run "f32 s$st nosync x150" float16384 150 MM_TCGEN05_STAGES=$st MM_TCGEN05_TILE_SYNC=0
done
for st in 4 5 6; do
  run "f32 s$st sync x10" float16384 10 MM_TCGEN05_STAGES=$st
done
for st in 4 5 6; do
  run "f16 s$st sync x30" half32768 30 MM_TCGEN05_STAGES=$st
done
run "f16 s4 nosync x30" half32768 30 MM_TCGEN05_STAGES=4 MM_TCGEN05_TILE_SYNC=0
