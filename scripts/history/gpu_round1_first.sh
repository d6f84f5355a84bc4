#!/bin/bash
# First GPU contact: each stage in its own process under a timeout so that a hung kernel cannot
# take the whole call down.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== semiring subset"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "semiring_shapes or distance" > gpurun_out/t_semiring.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_semiring.log
echo "== double"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "double" > gpurun_out/t_double.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_double.log
echo "== float tensor"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "float_tensor or tf32" > gpurun_out/t_float.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_float.log
echo "== half tensor"; timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "half" > gpurun_out/t_half.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_half.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_all.log
echo "== bench float4096"; timeout 300 python bench.py --workload float4096 --steps 5 > gpurun_out/bench_4096.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_4096.log
echo "== bench float16384"; timeout 900 python bench.py --steps 5 > gpurun_out/bench_16384.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_16384.log
echo "== bench double"; timeout 600 python bench.py --workload double8192 --steps 3 --no-e2e > gpurun_out/bench_double.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_double.log
echo "== bench addmin"; timeout 600 python bench.py --workload addmin8192 --steps 3 --no-e2e > gpurun_out/bench_addmin.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_addmin.log
