cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
R="--emulate-ranks 8 --steps 40 --no-e2e --no-cpu"
echo "== baseline overlap"; MM_DIAG_STAMPS=1 timeout 300 python bench.py $R 2>&1 | grep diag
echo "== grid 296"; MM_DIAG_STAMPS=1 MM_DIAG_PREP_GRID=296 timeout 300 python bench.py $R 2>&1 | grep diag
echo "== N=1"; MM_DIAG_STAMPS=1 timeout 300 python bench.py --steps 40 --no-e2e --no-cpu 2>&1 | grep diag
