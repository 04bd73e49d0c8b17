#!/bin/bash
O=gpurun_out/r2s2; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_shadow.py tests/test_gpu_i8.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -25) | tee $O/pytest.txt
for rep in 1 2; do for sh in 0 1; do echo "== 100M shadow=$sh"; DAWN_AB_SHADOW=$sh timeout 300 python tools/ab_gemm.py 100000000 1024 10 gemm_growth 0 2>&1 | tail -1; done; done | tee $O/ab_shadow_100m.txt
