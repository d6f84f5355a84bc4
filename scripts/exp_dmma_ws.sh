#!/bin/bash
# Ran at commit ad15438, which keeps every DMMA kernel generation behind MM_DMMA_TMA / MM_DMMA_WS /
# MM_DMMA_PRODUCERS / MM_DMMA_DEBUG; only the TMA-fed kernel remains afterwards.
set +e
mkdir -p gpurun_out/r01
echo "== double parity (TMA-fed kernel, default)"; timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "double or DOUBLE or golden" 2>&1 | tail -2 | tee gpurun_out/exp_dmma_ws3.log
timeout 600 python scripts/exp_dmma_ws.py 2>&1 | tee -a gpurun_out/exp_dmma_ws3.log
