#!/bin/bash
# Round-end measurements on 8 B200 of one box: BASELINE configs C3 (k=100), C4 (batch sweep) and C5 (int8).
set -u
O=gpurun_out/final8
mkdir -p $O
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 500 $TR 29511 bench.py --gpus $N --k 100 --no-cpu-baseline > $O/bench_100m_${N}gpu_k100_c3.json 2> $O/c3.err
timeout 500 $TR 29512 bench.py --gpus $N --rows 50000000 --no-cpu-baseline --steps 10 --latency-steps 100 \
    --sweep 1,2,3,4,8,16,32,64,128,256,512,1024,2048,4096 > $O/bench_50m_${N}gpu_c4_sweep.json 2> $O/c4.err
timeout 500 $TR 29513 bench.py --gpus $N --scalar i8 --rows 500000000 --batch 1 --steps 30 --warmup 5 --no-cpu-baseline \
    > $O/bench_i8_500m_${N}gpu_c5.json 2> $O/c5.err
timeout 500 $TR 29514 bench.py --gpus $N --no-cpu-baseline > $O/bench_100m_${N}gpu_k10.json 2> $O/k10.err
tail -c 600 $O/*.err
ls -la $O
