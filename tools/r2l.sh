#!/bin/bash
# micro-batching front under closed-loop load
O=gpurun_out/r2l; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
mkdir -p tools/bin
g++ -O2 -std=c++17 -I include tools/batcher_bench.cpp -o tools/bin/batcher_bench -L dawnsearch_b200/lib -ldawn_b200 -Wl,-rpath,$PWD/dawnsearch_b200/lib -lpthread || exit 1
nproc > $O/nproc.txt
(timeout 600 python -m pytest tests/test_gpu_front.py -m gpu -q -x 2>&1 | tail -5) | tee $O/pytest_front.txt
timeout 300 tools/bin/batcher_bench 10000000 10 3 1024 100 1,16,64,256,1024,2048 | tee $O/batcher_10m_v3.jsonl
timeout 300 tools/bin/batcher_bench 100000000 10 4 1024 100 64,1024,2048 | tee $O/batcher_100m_v3.jsonl
