#!/bin/bash
O=gpurun_out/r2m2; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_front.py tests/test_rust_shim.py -m gpu -q -x 2>&1 | tail -5) | tee $O/pytest_front.txt
