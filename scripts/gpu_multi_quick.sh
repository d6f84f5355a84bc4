#!/bin/bash
# bench.py under torchrun at N GPUs only (no single-GPU legs, no C++ driver): usage  bash scripts/gpu_multi_quick.sh <N>
set +e
N=${1:-2}
mkdir -p gpurun_out/r01
O=gpurun_out/r01
J='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print("%-12s n_gpus %d ms/step %.3f value %.0f | broadcast of B %.3f ms (%d bytes) | %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], c.get("broadcast_b_ms", -1), c.get("broadcast_b_bytes", 0), d["clocks"]["reasons"]))'
for wl in float16384 double8192; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $wl --gpus $N --steps 20 --no-cpu > $O/scale_${wl}_n$N.json 2>$O/scale_${wl}_n$N.err
  tail -1 $O/scale_${wl}_n$N.json | python -c "$J" "$wl" || tail -5 $O/scale_${wl}_n$N.err
done
