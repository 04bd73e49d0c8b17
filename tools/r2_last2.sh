#!/bin/bash
O=gpurun_out/r2_last2; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -4) | tee $O/pytest_2gpu.txt
