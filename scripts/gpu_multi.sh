cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_variants_gpu.py -q -m gpu -k "multi or default_entry" > gpurun_out/m${NG:-2}_tests.log 2>&1; echo "multi tests rc=$?"; tail -4 gpurun_out/m${NG:-2}_tests.log
timeout 600 python -m pytest tests/test_zz_host_exec_gpu.py -q -m gpu -k "mm_num_gpus" > gpurun_out/m${NG:-2}_host.log 2>&1; echo "host rc=$?"; tail -3 gpurun_out/m${NG:-2}_host.log
G=${NG:-2}
for wl in float16384 double8192; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 3 --workload $wl > gpurun_out/m${NG:-2}_bench_${wl}_n$G.log 2>&1; echo "bench $wl rc=$?"
  tail -1 gpurun_out/m${NG:-2}_bench_${wl}_n$G.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'N=',d['n_gpus'],'value',round(d['value']/1e3,1),'step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms'],3),'prep',round(d['roofline']['prep_ms'],3),'e2e',round(d['e2e']['value']/1e3,1), round(d['e2e']['ms_per_step'],2),'ms', d.get('broadcast_b'))
except Exception as e: print('parse failed', e)
"
done
bash scripts/build_host.sh /tmp/hb > /dev/null 2>&1
for g in 1 $G; do MM_NUM_GPUS=$g timeout 300 /tmp/hb/RunHardware 16384 16384 16384 hw off 2>&1 | grep -E "Kernel executed|failed" | tee -a gpurun_out/m${G}_runhardware.log; done
MM_NUM_GPUS=$G timeout 300 /tmp/hb/RunHardware 2048 2048 2048 hw on 2>&1 | grep -E "Kernel executed|verified|failed|Mismatch" | tee -a gpurun_out/m${G}_runhardware.log
