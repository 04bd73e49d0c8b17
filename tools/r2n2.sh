#!/bin/bash
# growth of the rounds, same box / same process A/B on one shard of C3 (12.5M rows)
O=gpurun_out/r2n2; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
{ echo "== k=100 (k'=128), gemm_growth 4 / 6 / 8"; timeout 200 python tools/ab_gemm.py 12500000 1024 100 gemm_growth 4,6,8 2>&1 | tail -3
  echo "== k=10 (k'=16), gemm_growth 16 / 32 / 64"; timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 16,32,64 2>&1 | tail -3
  echo "== k=20 (k'=32), gemm_growth 8 / 16 / 32"; timeout 200 python tools/ab_gemm.py 12500000 1024 20 gemm_growth 8,16,32 2>&1 | tail -3
  echo "== 100M rows, k=10, gemm_growth 16 / 32"; timeout 300 python tools/ab_gemm.py 100000000 1024 10 gemm_growth 16,32 2>&1 | tail -2
} | tee $O/ab_growth.txt
