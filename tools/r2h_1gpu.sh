#!/bin/bash
# staged-row selects + in-place load: full GPU suite, then the shard-size lines the change targets
set -u
O=gpurun_out/r2h
mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q -x) > $O/pytest_gpu.log 2>&1
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:100,1024:10 > $O/bench_12m5.json 2> $O/bench_12m5.err
timeout 300 python bench.py --scalar i8 --rows 62500000 --steps 10 --warmup 3 --no-cpu-baseline --sweep 1024:10,1024:100 > $O/bench_i8_62m5.json 2> $O/bench_i8.err
tail -5 $O/pytest_gpu.log
cat $O/bench_12m5.json $O/bench_i8_62m5.json | cut -c1-400
