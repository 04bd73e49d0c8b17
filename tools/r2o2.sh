#!/bin/bash
O=gpurun_out/r2o2; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_front.py tests/test_gpu_round2.py tests/test_gpu_sharded.py tests/test_rust_shim.py -m gpu -q -x 2>&1 | tail -15) | tee $O/pytest.txt
