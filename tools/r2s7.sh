#!/bin/bash
O=gpurun_out/r2s7; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
timeout 200 python bench.py --front multi --gpus 1 --rows 20000000 --steps 5 --warmup 3 --latency-steps 20 > $O/front_multi_1gpu_20m.json 2> $O/multi.err; tail -2 $O/multi.err; cut -c1-400 $O/front_multi_1gpu_20m.json; echo
for w in i8gemm shadow; do echo "== memcheck $w"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py $w 2>&1 | tail -3; done | tee $O/sanitizer_memcheck_two_stage_select.txt
