#!/bin/bash
# Multi-GPU evidence:  gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [TAG] [quick]'
# bench.py under torchrun at N GPUs (row-block split, one NCCL broadcast of B) for the float and double
# configurations, the single-GPU legs measured on the same box, and the native C++ driver (RunHardware with
# MM_NUM_GPUS: one host thread per GPU, ncclBroadcast of B).  `quick` = the torchrun legs only.
set +e
N=${1:-2}; TAG=${2:-r01}; MODE=${3:-full}
R=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd "$R"
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L | tee $O/multi_gpus_$N.txt
J='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print("%-12s n_gpus %d ms/step %.3f value %.0f | broadcast of B %s ms | %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], c.get("broadcast_b_ms"), d["clocks"]["reasons"]))'
for wl in float16384 double8192; do
  if [ $MODE != quick ]; then
    timeout 900 python bench.py --workload $wl --gpus 1 --steps 20 --no-cpu > $O/scale_${wl}_n1.json 2>/dev/null
    tail -1 $O/scale_${wl}_n1.json | python -c "$J" "$wl"
  fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --workload $wl --gpus $N --steps 20 --no-cpu > $O/scale_${wl}_n$N.json 2>$O/scale_${wl}_n$N.err
  tail -1 $O/scale_${wl}_n$N.json | python -c "$J" "$wl" || tail -5 $O/scale_${wl}_n$N.err
done
if [ $MODE != quick ]; then
  echo "== native C++ multi-GPU driver"
  bash scripts/build_host.sh /tmp/hostbuild > /dev/null 2>&1 || echo "host build failed"
  ( MM_NUM_GPUS=$N /tmp/hostbuild/RunHardware 1024 1024 1024 hw on; echo "rc=$?"
    MM_NUM_GPUS=$N /tmp/hostbuild/RunHardware 16384 16384 16384 hw off; echo "rc=$?" ) 2>&1 | tee $O/multi_runhardware_$N.log | tail -16
fi
