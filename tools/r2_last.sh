#!/bin/bash
O=gpurun_out/r2_last; mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-60
timeout 600 python -m pytest tests/test_gpu_shadow.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --latency-steps 20 --no-cpu-baseline > $O/bench_quick.json 2> $O/bench.err; cut -c1-160 $O/bench_quick.json
