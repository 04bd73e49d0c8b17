#!/bin/bash
# final-tree verification on ONE B200: build, GPU suite, smoke(), the driver's bench command, single-GPU lines of the configs
set -u
O=gpurun_out/r2_verify; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1; echo "build rc=$?" | tee -a $O/build.log
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log | head -2
(time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; tail -4 $O/smoke.log | head -1 | cut -c1-80
(time timeout 600 python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err; cut -c1-200 $O/bench_default.json
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:100,1024:10,1024:20,4096:10,256:10,1 > $O/bench_12m5.json 2> /dev/null
timeout 300 python bench.py --rows 10000000 --no-cpu-baseline --sweep 1,2,4,1:20,1:100 > $O/bench_10m_c2.json 2> /dev/null
