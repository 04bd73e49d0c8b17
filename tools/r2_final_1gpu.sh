#!/bin/bash
# Round-2 measurements on ONE B200 (files under gpurun_out/r2_final/; copied to profiles/ afterwards).
set -u
O=gpurun_out/r2_final
mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5) > $O/bench_100m_1gpu.json 2> $O/bench_100m_1gpu.err
(time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 300 python bench.py --rows 10000000 --no-cpu-baseline --sweep 1,2,4,1:20,1:100 > $O/bench_10m_1gpu_c2.json 2> /dev/null
timeout 300 python bench.py --scalar i8 --rows 62500000 --steps 10 --warmup 3 --no-cpu-baseline --sweep 1,2,16,256,1024:100,4096 > $O/bench_i8_62m5_1gpu.json 2> /dev/null
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:100,1024:10,4096:10,256:10 > $O/bench_12m5_shard_of_c3.json 2> /dev/null
timeout 300 python bench.py --rows 50000000 --no-cpu-baseline --steps 10 --latency-steps 100 --sweep 1,2,3,4,8,16,32,64,128,256,512,1024,2048,4096 > $O/bench_50m_1gpu_c4_sweep.json 2> /dev/null
timeout 120 python tools/small_latency.py > $O/small_corpus_latency.json 2> /dev/null
timeout 200 python tools/ingest_rate.py > $O/ingest_rate.json 2> /dev/null
timeout 400 python tools/c1_reference_path.py > $O/c1_reference_path.json 2> /dev/null
tail -4 $O/bench_100m_1gpu.err $O/bench_reference_arm.err
ls -la $O
