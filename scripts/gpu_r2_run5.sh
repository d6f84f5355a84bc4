cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # label, args...
  label=$1; shift
  timeout 600 python bench.py --no-e2e --no-cpu "$@" 2>&1 | tail -1 > /tmp/line.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open('/tmp/line.json'))
    print("%-52s step %.3f ms  kernel %.3f  prep %.3f" % (sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['prep_ms']))
except Exception as e:
    print(sys.argv[1], "FAILED", open('/tmp/line.json').read()[-300:])
PY
}
R="--emulate-ranks 8 --steps 200"
run "overlap (baseline)" $R
MM_DIAG_IGNORE_READY=1 run "overlap, GEMM ignores flags (WRONG results ok)" $R 
run "overlap cta_group=1" $R --tune cta_group=1
run "no overlap cta_group=1" $R --tune cta_group=1,b_overlap=0
MM_DIAG_PREP_GRID=296 run "overlap prep grid 296" $R
MM_DIAG_PREP_GRID=74 run "overlap prep grid 74" $R
MM_DIAG_PREP_SMEM=16000 run "overlap prep dyn smem 16000" $R
run "overlap stages=4 (GEMM smem 148K)" $R --tune stages=4
run "no overlap stages=4" $R --tune stages=4,b_overlap=0
run "overlap stages=3 (GEMM smem 115K)" $R --tune stages=3
run "no overlap stages=3" $R --tune stages=3,b_overlap=0
