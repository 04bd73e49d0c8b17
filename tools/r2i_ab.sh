#!/bin/bash
# A/B of the int8 select kernel (lib = rows staged in shared memory, lib_ab = thread-per-candidate reads from L2), one box
O=gpurun_out/r2i; mkdir -p $O
for rep in 1 2; do
  for lib in lib lib_ab; do
    for k in 10 100; do
      echo "== $lib rep $rep k$k"; DAWN_AB_SCALAR=i8 DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so timeout 200 python tools/ab_gemm.py 62500000 1024 $k gemm_growth 0 2>&1 | tail -1
    done
  done
done 2>&1 | tee $O/ab_select_i8.txt
