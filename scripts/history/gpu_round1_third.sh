#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== tensor subset, CTA pair (default)"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "float_tensor or half or tf32 or golden or transposed or lifecycle or row_block" > gpurun_out/t3_cg2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t3_cg2.log
echo "== tensor subset, single CTA"; MM_TCGEN05_CTA_GROUP=1 timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "float_tensor or half or tf32 or golden or transposed" > gpurun_out/t3_cg1.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t3_cg1.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t3_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t3_all.log
echo "== bench float16384 CG2"; timeout 900 python bench.py --steps 20 --no-cpu > gpurun_out/b3_f32_cg2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_f32_cg2.log | cut -c1-1500
echo "== bench float16384 CG1"; MM_TCGEN05_CTA_GROUP=1 timeout 900 python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/b3_f32_cg1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_f32_cg1.log | cut -c1-1200
echo "== bench float16384 CG2 100 steps"; timeout 900 python bench.py --steps 100 --no-cpu --no-e2e > gpurun_out/b3_f32_cg2_s100.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_f32_cg2_s100.log | cut -c1-1200
echo "== bench half32768 CG2"; timeout 900 python bench.py --workload half32768 --steps 5 --no-e2e --no-cpu > gpurun_out/b3_f16_cg2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_f16_cg2.log | cut -c1-1200
echo "== bench half32768 CG1"; MM_TCGEN05_CTA_GROUP=1 timeout 900 python bench.py --workload half32768 --steps 5 --no-e2e --no-cpu > gpurun_out/b3_f16_cg1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_f16_cg1.log | cut -c1-1200
echo "== bench double"; timeout 600 python bench.py --workload double8192 --steps 5 --no-cpu > gpurun_out/b3_double.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_double.log | cut -c1-1500
echo "== bench addmin"; timeout 600 python bench.py --workload addmin8192 --steps 5 --no-cpu > gpurun_out/b3_addmin.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b3_addmin.log | cut -c1-1500
echo "== ncu tf32 cg2"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof3_tcgen05_tf32_cg2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu3_tf32.log 2>&1; echo "rc=$?"
echo "== ncu addmin"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o gpurun_out/prof3_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu3_addmin.log 2>&1; echo "rc=$?"
echo "== ncu dmma"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -f -o gpurun_out/prof3_dmma python bench.py --workload double8192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu3_dmma.log 2>&1; echo "rc=$?"
