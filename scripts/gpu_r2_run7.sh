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
run "r8 overlap cache=0" $R
MM_DIAG_CACHE=1 run "r8 overlap cache=1 (ld.cs)" $R
MM_DIAG_CACHE=2 run "r8 overlap cache=2 (st.cs)" $R
MM_DIAG_CACHE=3 run "r8 overlap cache=3 (both)" $R
MM_DIAG_CACHE=3 MM_DIAG_PREP_GRID=74 run "r8 overlap cache=3 grid 74" $R
run "r8 no overlap" $R --tune b_overlap=0
R="--emulate-ranks 4 --steps 100"
run "r4 overlap cache=0" $R
MM_DIAG_CACHE=3 run "r4 overlap cache=3 (both)" $R
run "r4 no overlap" $R --tune b_overlap=0
MM_DIAG_CACHE=3 run "N=1 overlap cache=3 (both)" --steps 50
run "N=1 no overlap" --steps 50 --tune b_overlap=0
