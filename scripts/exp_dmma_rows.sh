#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== double parity (auto tile height)"; timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "double or DOUBLE or golden" 2>&1 | tail -2 | tee gpurun_out/exp_dmma_rows.log
echo "== double parity (MM_DMMA_WS=0: all warps prefetch)"; MM_DMMA_WS=0 timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "double or DOUBLE or golden" 2>&1 | tail -2 | tee -a gpurun_out/exp_dmma_rows.log
echo "== double parity (forced 64-row tiles)"; MM_DMMA_TILE_ROWS=64 timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "double or DOUBLE or golden" 2>&1 | tail -2 | tee -a gpurun_out/exp_dmma_rows.log
timeout 900 python scripts/exp_dmma_rows.py 2>&1 | tee -a gpurun_out/exp_dmma_rows.log
