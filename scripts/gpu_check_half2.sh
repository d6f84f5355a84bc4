cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -q -m gpu -k "HALF or half or golden" > gpurun_out/h2_tests.log 2>&1; tail -2 gpurun_out/h2_tests.log
timeout 300 python -m pytest tests/test_zz_host_exec_gpu.py -q -m gpu -k half > gpurun_out/h2_host.log 2>&1; tail -1 gpurun_out/h2_host.log
for lib in before new; do
  if [ $lib = before ]; then export MM_B200_LIB=$GRAFT_REPO_ROOT/gemm_hls_b200/exp_libs/libmm_b200_before_half2.so; else unset MM_B200_LIB; fi
  timeout 300 python bench.py --workload half8192 --flags 2 --steps 5 --no-e2e --no-cpu > gpurun_out/h2_bench_$lib.log 2>&1
  tail -1 gpurun_out/h2_bench_$lib.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$lib', 'half EXACT 8192^3:', round(d['value']/1e3,2),'TOp/s', round(d['ms_per_step'],2),'ms')
except Exception as e: print('$lib failed', e)"
  tail -3 gpurun_out/h2_bench_$lib.log | cut -c1-300
done
