#!/bin/bash
# final-tree verification on ONE B200: build, GPU suite, smoke(), the driver's two bench commands
set -u
O=gpurun_out/r2_verify; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1; echo "build rc=$?" | tee -a $O/build.log
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
(time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; tail -5 $O/smoke.log
(time timeout 600 python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err; cut -c1-600 $O/bench_default.json
(time timeout 900 python bench.py --impl reference) > $O/bench_reference.json 2> $O/bench_reference.err; tail -3 $O/bench_reference.err; cut -c1-400 $O/bench_reference.json
timeout 300 python bench.py --rows 12500000 --k 20 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:20,1024:10,1024:100 > $O/bench_12m5_k20.json 2> /dev/null
