mkdir -p gpurun_out/r2k
(time timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x --durations=5) 2>&1 | tail -25 | tee gpurun_out/r2k/pytest_full_size.txt
