# Round-2 closing run on one B200: full suite, smoke, default bench line + reference arm, ncu of the kind::i8 kernel.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_ref.json; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -o gpurun_out/r02_gemm_i8 -f python bench.py --workload uint8_16384 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_i8.log 2>&1; echo "ncu i8 rc=$?"
