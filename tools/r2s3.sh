#!/bin/bash
O=gpurun_out/r2s3; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
timeout 200 python tools/shadow_batch1.py 10000000 | tee $O/shadow_batch1_10m.json
timeout 300 python tools/shadow_batch1.py 100000000 | tee $O/shadow_batch1_100m.json
(time timeout 600 python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err; cut -c1-300 $O/bench_default.json
