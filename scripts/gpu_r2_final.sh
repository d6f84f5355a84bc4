# Round-2 closing run on one B200: sanitizer passes, ncu of the (Add, Min) ring kernel, full suite, smoke, default bench.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_sanitize.sh r02 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:semiring_ring -s 1 -c 1 -o gpurun_out/r02_semiring_addmin -f python bench.py --workload addmin8192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_semiring.log 2>&1; echo "ncu semiring rc=$?"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench.json
