# Full GPU suite + default bench + ncu evidence for the tcgen05 GEMM (one B200).  Outputs under gpurun_out/.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/full_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/full_bench.err | tail -1 > gpurun_out/full_bench.json; cut -c1-400 gpurun_out/full_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>>gpurun_out/full_bench.err | tail -1 > gpurun_out/full_bench_ref.json; cut -c1-300 gpurun_out/full_bench_ref.json
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_float16384.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# one full capture of the GEMM kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -o gpurun_out/r02_gemm_tf32 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
