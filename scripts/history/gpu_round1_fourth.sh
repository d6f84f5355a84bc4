#!/bin/bash
set +e
mkdir -p gpurun_out
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(sys.argv[1], "ms/step %.3f value %.0f | kernel_ms %.3f prep_ms %.3f achieved %.1f frac %.3f | e2e %s | clocks %s" % (d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], r["achieved"], r["frac"], (d.get("e2e") or {}).get("value"), d["clocks"]))'
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t4_all.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t4_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 > gpurun_out/b4_f32.log 2>&1; tail -1 gpurun_out/b4_f32.log | python -c "$J" f32_cg2_s20
timeout 900 python bench.py --steps 100 --no-e2e --no-cpu > gpurun_out/b4_f32_s100.log 2>&1; tail -1 gpurun_out/b4_f32_s100.log | python -c "$J" f32_cg2_s100
MM_TCGEN05_CTA_GROUP=1 timeout 900 python bench.py --steps 100 --no-e2e --no-cpu > gpurun_out/b4_f32_s100_cg1.log 2>&1; tail -1 gpurun_out/b4_f32_s100_cg1.log | python -c "$J" f32_cg1_s100
timeout 900 python bench.py --workload half32768 --steps 20 --no-e2e --no-cpu > gpurun_out/b4_f16_cg2.log 2>&1; tail -1 gpurun_out/b4_f16_cg2.log | python -c "$J" f16_cg2_s20
MM_TCGEN05_CTA_GROUP=1 timeout 900 python bench.py --workload half32768 --steps 20 --no-e2e --no-cpu > gpurun_out/b4_f16_cg1.log 2>&1; tail -1 gpurun_out/b4_f16_cg1.log | python -c "$J" f16_cg1_s20
timeout 900 python bench.py --workload addmin8192 --steps 10 --no-e2e --no-cpu > gpurun_out/b4_addmin.log 2>&1; tail -1 gpurun_out/b4_addmin.log | python -c "$J" addmin
timeout 900 python bench.py --workload double8192 --steps 10 --no-e2e --no-cpu > gpurun_out/b4_double.log 2>&1; tail -1 gpurun_out/b4_double.log | python -c "$J" double
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/b4_reference.log 2>&1; tail -1 gpurun_out/b4_reference.log | cut -c1-400
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches4_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; grep -c "gemm_tcgen05" gpurun_out/launches4_float16384.csv
echo "== ncu full tf32 cg2"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof4_tcgen05_tf32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== ncu full f16 cg2"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof4_tcgen05_f16 python bench.py --workload half32768 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== ncu addmin"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o gpurun_out/prof4_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
echo "== ncu dmma"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -f -o gpurun_out/prof4_dmma python bench.py --workload double8192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; echo "rc=$?"
