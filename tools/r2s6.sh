#!/bin/bash
O=gpurun_out/r2s6; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_shadow.py tests/test_gpu_i8.py -m gpu -q -x 2>&1 | tail -3) | tee $O/pytest.txt
DAWN_OPTS=shadow_i8=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_12m5_b1024_k10_lib.csv python tools/ncu_target.py f16gemm 12500000 1024 10 > $O/ncu_lib.log 2>&1
for lib in lib lib_ab; do for k in 10 20; do
  echo "== $lib k$k"; DAWN_AB_SHADOW=1 DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so timeout 200 python tools/ab_gemm.py 12500000 1024 $k gemm_growth 0 2>&1 | tail -1
done; done | tee $O/ab_select.txt
