cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python scripts/diag_tcgen05.py > gpurun_out/r2_diag.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r2_diag.log
