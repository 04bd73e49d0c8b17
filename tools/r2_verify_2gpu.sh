#!/bin/bash
# final-tree verification on TWO B200: the multi-GPU tests (NCCL, dawn_multi, batcher over dawn_multi) and the driver's N=2 commands
O=gpurun_out/r2_verify2; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py tests/test_gpu_front.py -m gpu -q 2>&1 | tail -6) | tee $O/pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29531 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $O/reference_2gpu.json 2> $O/reference_2gpu.err; cut -c1-200 $O/reference_2gpu.json
timeout 400 $TR 29532 bench.py --gpus 2 > $O/headline_k10_2gpu.json 2> $O/bench_2gpu.err; cut -c1-300 $O/headline_k10_2gpu.json; echo
timeout 400 python bench.py --front multi --gpus 2 --steps 10 --warmup 3 --latency-steps 100 > $O/front_multi_k10_2gpu.json 2> $O/multi.err; cut -c1-300 $O/front_multi_k10_2gpu.json; echo
