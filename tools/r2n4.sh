#!/bin/bash
O=gpurun_out/r2n4; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
{ echo "== k=50 (k'=64), gemm_growth 8 / 4"; timeout 200 python tools/ab_gemm.py 12500000 1024 50 gemm_growth 8,4 2>&1 | tail -2
  echo "== k=50 (k'=64), gemm_growth 4 / 8"; timeout 200 python tools/ab_gemm.py 12500000 1024 50 gemm_growth 4,8 2>&1 | tail -2
  echo "== k=20 (k'=32), gemm_growth 16 / 8"; timeout 200 python tools/ab_gemm.py 12500000 1024 20 gemm_growth 16,8 2>&1 | tail -2
} | tee $O/ab_growth3.txt
