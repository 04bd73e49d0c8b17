#!/bin/bash
# 8 B200, int8 shadow on (bench default): the headline config and C3
O=gpurun_out/r2_8gpu_shadow; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29541 bench.py --gpus 8 --no-cpu-baseline > $O/headline_k10_8gpu.json 2> $O/headline.err; cut -c1-250 $O/headline_k10_8gpu.json; echo
timeout 300 $TR 29542 bench.py --gpus 8 --k 100 --no-cpu-baseline --latency-steps 50 > $O/c3_k100_8gpu.json 2> $O/c3.err; cut -c1-250 $O/c3_k100_8gpu.json; echo
