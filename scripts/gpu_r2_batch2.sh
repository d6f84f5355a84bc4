cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
./scripts/_build/exp_pipe_rates | tee gpurun_out/r02_exp_pipe_rates.jsonl
timeout 1200 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py -q -m gpu -x > gpurun_out/b2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/b2_tests.log
for wl in uint8_16384 addmin8192; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>gpurun_out/b2_bench_$wl.err | tail -1 > gpurun_out/b2_bench_$wl.json
python - $wl <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/b2_bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value']/1e3,1), 'T/s step', round(d['ms_per_step'],3), 'roof', d['roofline']['frac'], 'e2e', round(d['e2e']['value']/1e3,1), d.get('check'))
except Exception as e: print('bench failed', sys.argv[1], e, open('gpurun_out/b2_bench_%s.err'%sys.argv[1]).read()[-500:])
PY
done
timeout 1500 python scripts/tile_sweep.py --workload half32768 --steps 5 --out gpurun_out/r02_tile_sweep_half32768.csv 2>&1 | tail -30
