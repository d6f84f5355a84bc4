# One torchrun bench line at NG GPUs (default workload float16384): bash scripts/gpu_multi_short.sh [workload ...]
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
G=${NG:-8}
for wl in ${@:-float16384}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 3 --workload $wl > gpurun_out/ms${G}_bench_${wl}.log 2>&1; echo "bench $wl rc=$?"
  tail -1 gpurun_out/ms${G}_bench_${wl}.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'N=',d['n_gpus'],'value',round(d['value']/1e3,1),'step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms'],3),'prep',round(d['roofline']['prep_ms'],3),'e2e',round(d['e2e']['value']/1e3,1), round(d['e2e']['ms_per_step'],2),'ms', d['e2e'].get('host_memory'))
except Exception as e: print('parse failed', e); print(open('gpurun_out/ms${G}_bench_${wl}.log').read()[-1500:])
"
done
