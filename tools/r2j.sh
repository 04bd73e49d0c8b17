mkdir -p gpurun_out/r2j
timeout 300 python tools/f3_bound.py 12500000 10 | tee gpurun_out/r2j/f3_bound_k10.json
timeout 300 python tools/f3_bound.py 12500000 100 | tee gpurun_out/r2j/f3_bound_k100.json
