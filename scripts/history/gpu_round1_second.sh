#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_all2.log
echo "== tf32 rounding experiment"; timeout 300 python scripts/exp_tf32_rounding.py > gpurun_out/exp_tf32_rounding.json 2>&1; cat gpurun_out/exp_tf32_rounding.json
echo "== bench float16384 (100 steps, sustained)"; timeout 900 python bench.py --steps 100 > gpurun_out/bench_16384_s100.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_16384_s100.log
echo "== bench float16384 (10 steps)"; timeout 900 python bench.py --steps 10 --no-cpu > gpurun_out/bench_16384_s10.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_16384_s10.log
echo "== bench half32768"; timeout 900 python bench.py --workload half32768 --steps 5 --no-e2e --no-cpu > gpurun_out/bench_half32768.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_half32768.log
echo "== library baselines"; timeout 600 python scripts/library_baselines.py > gpurun_out/library_baselines.json 2>&1; cat gpurun_out/library_baselines.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/launches_float16384.csv
echo "== ncu full tf32"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof_tcgen05_tf32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_tf32.log 2>&1; echo "rc=$?"
echo "== ncu full prep"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"transpose_prep|round_tf32" -s 2 -c 2 -f -o gpurun_out/prof_prep python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_prep.log 2>&1; echo "rc=$?"
echo "== ncu full dmma"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -f -o gpurun_out/prof_dmma python bench.py --workload double8192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_dmma.log 2>&1; echo "rc=$?"
echo "== ncu full semiring addmin"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:semiring_tile -s 1 -c 1 -f -o gpurun_out/prof_semiring_addmin python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_addmin.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
