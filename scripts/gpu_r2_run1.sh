set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
# quick smoke of the new default path first (fail fast)
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r1_smoke.log
timeout 900 python -m pytest tests/test_variants_gpu.py -q -m gpu > gpurun_out/r1_variants.log 2>&1; echo "variants rc=$?"
tail -5 gpurun_out/r1_variants.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_zz_host_exec_gpu.py -q -m gpu > gpurun_out/r1_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r1_parity.log
for t in "" "b_overlap=0" "tma_store=0" "b_mn=0" "b_mn=0,tma_store=0" "block_n=128"; do
  echo "== tune: $t"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --tune "$t" 2>&1 | tail -1 | tee -a gpurun_out/r1_bench_tune.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['prep_ms'], d['clocks']['sm_mhz'], d['clocks']['samples'])"
done
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r1_bench_full.json; cat gpurun_out/r1_bench_full.json | cut -c1-1500
