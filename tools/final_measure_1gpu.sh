#!/bin/bash
# Round-end measurements on ONE B200 (every line a file under gpurun_out/final/; copied to profiles/ afterwards).
set -u
O=gpurun_out/final
mkdir -p $O tools/bin
[ -x tools/bin/scan_trace ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDAWN_SCAN_TRACE -I include tools/scan_trace.cu -o tools/bin/scan_trace
timeout 600 python bench.py                                   > $O/bench_100m_1gpu.json      2> $O/bench_100m_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 300 python bench.py --rows 10000000 --no-cpu-baseline --sweep 1,2,4,1:20,1:100 > $O/bench_10m_1gpu_c2.json 2> /dev/null
timeout 300 python bench.py --scalar i8 --rows 62500000 --batch 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_i8_62m5_1gpu_batch1.json 2> /dev/null
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:100,1024:10,4096:10,256:10 > $O/bench_12m5_shard_of_c3.json 2> /dev/null
timeout 120 python tools/small_latency.py                      > $O/small_corpus_latency.json 2> /dev/null
(timeout 60 tools/bin/scan_trace 100000 32; timeout 60 tools/bin/scan_trace 1000000 32; timeout 60 tools/bin/scan_trace 10000000 16) > $O/scan_timeline.txt 2>&1
timeout 120 python tools/ingest_rate.py                        > $O/ingest_rate_2m.json 2> /dev/null
timeout 400 python tools/c1_reference_path.py                  > $O/c1_reference_path_hnsw_recall.json 2> /dev/null
ls -la $O
