#!/bin/bash
# Round-end check of the final tree on ONE B200: full GPU test suite, default bench line, int8 shard, ncu evidence.
set -u
O=gpurun_out/final
mkdir -p $O tools/bin
[ -x tools/bin/scan_trace ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDAWN_SCAN_TRACE -I include tools/scan_trace.cu -o tools/bin/scan_trace
(timeout 600 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_100m_1gpu.json 2> $O/bench_100m_1gpu.err
timeout 300 python bench.py --scalar i8 --rows 62500000 --batch 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_i8_62m5_1gpu_batch1.json 2> /dev/null
timeout 120 python tools/small_latency.py > $O/small_corpus_latency.json 2> /dev/null
(timeout 60 tools/bin/scan_trace 100000 32; timeout 60 tools/bin/scan_trace 1000000 32; timeout 60 tools/bin/scan_trace 10000000 16) > $O/scan_timeline.txt 2>&1
B1="python bench.py --rows 10000000 --batch 1 --steps 40 --warmup 5 --latency-steps 0 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_topk_f16 --launch-skip 30 -c 1 -f -o $O/prof_scan_q1_10m $B1 > /dev/null 2>&1
ncu -i $O/prof_scan_q1_10m.ncu-rep --page raw --csv > $O/scan_q1_10m_ncu_full_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O | tail -12
