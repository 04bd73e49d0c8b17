#!/bin/bash
# the first batcher (one worker, global condition variables; lib_ab) under the same load generator as the rewrite
O=gpurun_out/r2l; mkdir -p $O tools/bin
g++ -O2 -std=c++17 -I include tools/batcher_bench.cpp -o tools/bin/batcher_bench_v1 -L dawnsearch_b200/lib_ab -ldawn_b200 -Wl,-rpath,$PWD/dawnsearch_b200/lib_ab -lpthread || exit 1
timeout 300 tools/bin/batcher_bench_v1 10000000 10 3 1024 100 1,16,64,256,1024,2048 | tee $O/batcher_10m_v1_same_bench.jsonl
timeout 300 tools/bin/batcher_bench_v1 100000000 10 4 1024 100 64,2048 | tee $O/batcher_100m_v1_same_bench.jsonl
