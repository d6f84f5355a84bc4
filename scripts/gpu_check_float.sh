cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -q -m gpu -k "float or tf32 or golden or tuning or multi or lifecycle or graph" 2>&1 | tail -2
for r in 1 8; do
timeout 300 python bench.py --steps 50 --no-e2e --no-cpu --emulate-ranks $r 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ranks', $r, 'step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'prep', round(d['roofline']['prep_ms'],3), 'value', round(d['value']/1e3,1))"
done
