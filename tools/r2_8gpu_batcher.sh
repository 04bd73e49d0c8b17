#!/bin/bash
# the complete one-process serving path on 8 B200: single-query callers -> dawn_batcher -> dawn_multi (8 shards, NCCL) -> merge
O=gpurun_out/r2_8gpu_batcher; mkdir -p $O tools/bin
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
g++ -O2 -std=c++17 -I include tools/batcher_bench.cpp -o tools/bin/batcher_bench -L dawnsearch_b200/lib -ldawn_b200 -Wl,-rpath,$PWD/dawnsearch_b200/lib -lpthread || exit 1
nvidia-smi -L | wc -l > $O/gpus.txt; nproc > $O/nproc.txt
timeout 400 tools/bin/batcher_bench 100000000 10 3 1024 100 1,64,256,1024,2048,4096 8 2> $O/batcher_100m_8gpu.err | tee $O/batcher_100m_8gpu.jsonl
tail -3 $O/batcher_100m_8gpu.err
