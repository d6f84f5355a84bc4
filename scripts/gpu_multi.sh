#!/bin/bash
# Multi-GPU evidence: bench.py under torchrun at N GPUs (row-block split, one NCCL broadcast of B),
# the 2-rank invariant test, and the native C++ driver (RunHardware.exe with MM_NUM_GPUS).
# usage: bash scripts/gpu_multi.sh <N>
set +e
N=${1:-2}
mkdir -p gpurun_out/r01
O=gpurun_out/r01
nvidia-smi -L | tee $O/multi_gpus_$N.txt
J='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; c=d["clocks"]; print("%-24s n_gpus %d ms/step %.3f value %.0f | rank0 kernel_ms %.3f prep_ms %.3f | e2e %s | sm_mhz %s %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], r["kernel_ms"], r["prep_ms"], (d.get("e2e") or {}).get("value"), c["sm_mhz"], c["reasons"]))'
for wl in float16384 double8192; do
  for n in 1 $N; do
    if [ $n = 1 ]; then
      timeout 900 python bench.py --workload $wl --gpus 1 --steps 20 --no-cpu > $O/scale_${wl}_n1.json 2>/dev/null
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $wl --gpus $n --steps 20 --no-cpu > $O/scale_${wl}_n$n.json 2>$O/scale_${wl}_n$n.err
    fi
    tail -1 $O/scale_${wl}_n$n.json | python -c "$J" "$wl"
  done
done
echo "== native C++ multi-GPU driver"
mkdir -p /tmp/hostbuild && cd /tmp/hostbuild && python - <<'PY'
import re,os
root=os.environ.get("GRAFT_REPO_ROOT","/root/repo")
t=open(root+"/gemm_hls_b200/host/Config.h.in").read()
cfg=dict(MM_HOST_DATA_TYPE="float",MM_DATA_TYPE="float",MM_DTYPE_CODE="MM_DTYPE_FLOAT",MM_MAP_OP_UPPER="MULTIPLY",MM_MAP_OP="Multiply",MM_REDUCE_OP_UPPER="ADD",MM_REDUCE_OP="Add",MM_MEMORY_BUS_WIDTH_K=64,MM_MEMORY_BUS_WIDTH_M=64,MM_SIZE_N=512,MM_SIZE_K=512,MM_SIZE_M=512,MM_MEMORY_TILE_SIZE_N=128,MM_MEMORY_TILE_SIZE_M=256)
t=re.sub(r"\$\{(\w+)\}",lambda m:str(cfg[m.group(1)]),t).replace("#cmakedefine MM_EXACT","/* #undef MM_EXACT */")
open("Config.h","w").write(t)
PY
R=${GRAFT_REPO_ROOT:-/root/repo}
g++ -std=c++17 -O2 -DMM_DYNAMIC_SIZES -DMM_HAS_NCCL -I. -I$R/include -I$R/gemm_hls_b200/host -I/usr/local/cuda/include $R/gemm_hls_b200/host/RunHardware.cpp -L$R/gemm_hls_b200 -lmm_b200 -Wl,-rpath,$R/gemm_hls_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -lpthread -o RunHardware || echo "build failed"
cd $R
( MM_NUM_GPUS=$N /tmp/hostbuild/RunHardware 1024 1024 1024 hw on; echo "rc=$?"; MM_NUM_GPUS=$N /tmp/hostbuild/RunHardware 16384 16384 16384 hw off; echo "rc=$?" ) 2>&1 | tee $O/multi_runhardware_$N.log | tail -16
