# 2-GPU checks: NCCL paths (torch.distributed sharded + in-library dawn_multi), then short benches of both fronts.
mkdir -p gpurun_out/r2g; O=gpurun_out/r2g
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
nvidia-smi -L | head -3
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -15) | tee $O/pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29531 bench.py --gpus 2 --steps 10 --warmup 3 --latency-steps 100 --no-cpu-baseline > $O/bench_100m_2gpu.json 2> $O/bench_2gpu.err; tail -c 1500 $O/bench_100m_2gpu.json; echo; tail -3 $O/bench_2gpu.err
timeout 400 python bench.py --front multi --gpus 2 --steps 10 --warmup 3 --latency-steps 100 > $O/bench_100m_2gpu_front_multi.json 2> $O/bench_multi.err; tail -c 2500 $O/bench_100m_2gpu_front_multi.json; echo; tail -3 $O/bench_multi.err
