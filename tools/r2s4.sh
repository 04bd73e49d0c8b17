#!/bin/bash
O=gpurun_out/r2s4; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_shadow.py tests/test_gpu_i8.py tests/test_gpu_full_size.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -15) | tee $O/pytest.txt
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --no-parity-check --steps 20 --warmup 5 --sweep 1024:100,1024:10,1024:20,256:10,1 > $O/bench_12m5.json 2> /dev/null
DAWN_AB_SHADOW=1 timeout 200 python tools/ab_gemm.py 12500000 1024 100 shadow_big_k_rows 0,40000000 2>&1 | tail -2 | tee $O/ab_k100_12m5.txt
timeout 300 python bench.py --scalar i8 --rows 62500000 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --sweep 1024:10,1024:100 > $O/bench_i8_62m5.json 2> /dev/null
DAWN_AB_SHADOW=1 timeout 300 python tools/ab_gemm.py 100000000 1024 10 gemm_growth 0 2>&1 | tail -1 | tee $O/ab_100m.txt
