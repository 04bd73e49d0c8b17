#!/bin/bash
# C4 (batch sweep at 50M rows) and the headline config on 4 B200.
set -u
N=4; O=gpurun_out/r2_4gpu; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29513 bench.py --gpus $N --rows 50000000 --no-cpu-baseline --steps 10 --latency-steps 100 \
    --sweep 1,2,4,8,16,64,128,256,1024,4096 > $O/c4_sweep_50m.json 2> $O/c4.err
timeout 400 $TR 29514 bench.py --gpus $N --no-cpu-baseline > $O/headline_k10.json 2> $O/k10.err
tail -c 200 $O/c4.err; ls -la $O
