#!/bin/bash
# Diagnostics: stall breakdown of the tcgen05 kernel, single CTA vs CTA pair
set +e
mkdir -p gpurun_out
echo "== correctness CG2"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "float_tensor or half or tf32 or golden or transposed or row_block" 2>&1 | tail -2
for cg in 1 2; do
  echo "== CG=$cg"
  MM_TCGEN05_CTA_GROUP=$cg MM_TCGEN05_DEBUG=1 timeout 300 python bench.py --workload float16384 --steps 2 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "tcgen05 debug|kernel_ms" | cut -c1-300 | tail -2
  MM_TCGEN05_CTA_GROUP=$cg timeout 300 python bench.py --workload float16384 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('f32 cg$cg steps20: ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'TF', d['roofline']['achieved'], d['clocks'])"
  MM_TCGEN05_CTA_GROUP=$cg timeout 300 python bench.py --workload half32768 --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('f16 cg$cg steps5: ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'TF', d['roofline']['achieved'], d['clocks'])"
done 2>&1 | tee gpurun_out/exp_cta_group.log
