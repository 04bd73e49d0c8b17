#!/bin/bash
O=gpurun_out/r2_last3; mkdir -p $O
timeout 300 python bench.py --no-shadow --no-cpu-baseline > $O/bench_no_shadow.json 2> $O/bench.err; cut -c1-200 $O/bench_no_shadow.json
