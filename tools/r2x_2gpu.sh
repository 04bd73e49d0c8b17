mkdir -p gpurun_out/r2x; O=gpurun_out/r2x
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -6) | tee $O/pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29531 bench.py --gpus 2 --no-cpu-baseline > $O/headline_k10_2gpu.json 2> $O/bench_2gpu.err; tail -c 300 $O/headline_k10_2gpu.json; echo
timeout 400 python bench.py --front multi --gpus 2 --steps 10 --warmup 3 --latency-steps 100 > $O/front_multi_k10_2gpu.json 2> $O/multi.err; tail -c 200 $O/front_multi_k10_2gpu.json; echo
# racecheck on the tensor-core kernels (small shapes), for the record
for w in gemm i8gemm; do echo "== racecheck $w"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_smoke.py $w 2>&1 | grep -v "Host Frame\|^=========$" | tail -12; done > $O/sanitizer_racecheck.txt 2>&1
tail -30 $O/sanitizer_racecheck.txt
