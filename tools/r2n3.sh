#!/bin/bash
O=gpurun_out/r2n3; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
{ echo "== k=20 (k'=32), gemm_growth 32 / 16 / 8 / 4 (reversed order)"; timeout 200 python tools/ab_gemm.py 12500000 1024 20 gemm_growth 32,16,8,4 2>&1 | tail -4
  echo "== k=10 (k'=16), gemm_growth 32 / 16 / 8 (reversed order)"; timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 32,16,8 2>&1 | tail -3
  echo "== k=20, 100M rows, gemm_growth 16 / 8"; timeout 300 python tools/ab_gemm.py 100000000 1024 20 gemm_growth 16,8 2>&1 | tail -2
} | tee $O/ab_growth2.txt
