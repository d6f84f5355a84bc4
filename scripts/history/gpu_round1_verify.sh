#!/bin/bash
# End-of-round check of HEAD on one B200: smoke(), the default bench line, the reference arm, the semiring bench
set +e
mkdir -p gpurun_out/r01
O=gpurun_out/r01
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== default bench"; timeout 600 python bench.py > $O/bench_float16384_default.json 2>$O/bench_float16384_default.err; tail -1 $O/bench_float16384_default.json | cut -c1-1500
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>/dev/null; tail -1 $O/bench_reference_arm.json | cut -c1-400
echo "== addmin"; timeout 300 python bench.py --workload addmin8192 --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-1200
