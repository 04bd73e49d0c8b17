#!/bin/bash
# select_i8_kernel: one-stage (lib_ab, thread-per-candidate exact re-score) vs two-stage (lib), per-launch times under ncu + same-box A/B
O=gpurun_out/r2s5; mkdir -p $O
for lib in lib lib_ab; do
  DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so DAWN_OPTS=shadow_i8=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_12m5_b1024_k10_$lib.csv python tools/ncu_target.py f16gemm 12500000 1024 10 > $O/ncu_$lib.log 2>&1
done
for rep in 1 2; do for lib in lib lib_ab; do for k in 10 20; do
  echo "== $lib rep $rep k$k"; DAWN_AB_SHADOW=1 DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so timeout 200 python tools/ab_gemm.py 12500000 1024 $k gemm_growth 0 2>&1 | tail -1
done; done; done | tee $O/ab_select.txt
