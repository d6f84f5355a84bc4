#!/bin/bash
# A/B libraries for the float semiring kernel (experiments; the product is gemm_hls_b200/libmm_b200.so):
#   swap      scripts/next/semiring_swapped_pair.patch   second-k pair swapped (register parity of the FMNMX3 sources)
#   ring      scripts/next/semiring_tma_ring.patch       A and B tiles by TMA into a 4-stage ring, no block-wide barrier
#   swapring  both
# Each variant = a patched copy of csrc/ in /tmp, the seven float instantiation units recompiled, linked with the
# product's other objects into gemm_hls_b200/exp_libs/libmm_b200_<variant>.so (git-ignored, travels to the GPU box).
# Select at run time with MM_B200_LIB=<path> (gemm_hls_b200/__init__.py).
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
python $R/gemm_hls_b200/build.py > /dev/null
OUT=$R/gemm_hls_b200/exp_libs; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for v in ${@:-swap ring swapring}; do
  T=/tmp/semvar/$v; rm -rf $T; mkdir -p $T/gemm_hls_b200; cp -r $R/gemm_hls_b200/csrc $T/gemm_hls_b200/; cp -r $R/include $T/
  case $v in
    swap) (cd $T && patch -s -p1 < $R/scripts/next/semiring_swapped_pair.patch);;
    ring) (cd $T && patch -s -p1 < $R/scripts/next/semiring_tma_ring.patch);;
    swapring) (cd $T && patch -s -p1 < $R/scripts/next/semiring_tma_ring.patch && patch -s -p1 < $R/scripts/next/semiring_swapped_pair.patch);;
  esac
  objs=""
  for mp in 0 1 2 3 4 5 6; do
    nvcc $FLAGS -DMM_INST_T=float -DMM_INST_MAP=$mp -c $T/gemm_hls_b200/csrc/semiring_inst.cu -o $T/semiring_f32_$mp.o &
    objs="$objs $T/semiring_f32_$mp.o"
  done
  wait
  others=$(ls $R/gemm_hls_b200/build/*.o | grep -v semiring_f32_)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libmm_b200_$v.so $others $objs -Xlinker --exclude-libs=ALL
  echo "built $OUT/libmm_b200_$v.so"
done
