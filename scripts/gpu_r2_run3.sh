cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants_gpu.py -q -m gpu -k "multi or default_entry" > gpurun_out/r3_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/r3_multi.log
run() { # label, args...
  label=$1; shift
  timeout 600 python bench.py --no-e2e --no-cpu "$@" 2>&1 | tail -1 > /tmp/line.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open('/tmp/line.json'))
    print("%-44s value %9.1f  step %.3f ms  kernel %.3f  prep %.3f  sm %s MHz  P %s W" % (sys.argv[1], d['value']/1e3, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['prep_ms'], d['clocks']['sm_mhz'], round(d['clocks']['power_w_avg_under_load'] or 0)))
    d['label']=sys.argv[1]; open('gpurun_out/r3_bench.jsonl','a').write(json.dumps(d)+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", open('/tmp/line.json').read()[-300:])
PY
}
for r in 8 4 2; do
  for t in "" "b_overlap=0" "b_mn=0" "b_mn=0,tma_store=0"; do
    run "ranks=$r tune=[$t]" --emulate-ranks $r --steps 200 --tune "$t"
  done
done
for t in "" "b_overlap=0" "b_mn=0" "b_mn=0,tma_store=0"; do
  run "N=1 x100 tune=[$t]" --steps 100 --tune "$t"
done
for t in "" "tma_store=0"; do
  run "half32768 x20 tune=[$t]" --workload half32768 --steps 20 --tune "$t"
done
run "double8192 x10" --workload double8192 --steps 10
run "addmin8192 x10" --workload addmin8192 --steps 10
