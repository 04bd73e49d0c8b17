mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15) | tee $O/pytest.txt
timeout 200 python tools/i8_tensor_bench.py 62500000 1024 1 > $O/i8_bench.json 2> $O/i8_bench.err; cat $O/i8_bench.json; tail -3 $O/i8_bench.err
for rep in 1 2 3; do
  for lib in lib lib_ab; do
    echo "== $lib rep $rep k10"; DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 0 2>&1 | tail -1
    echo "== $lib rep $rep k100"; DAWN_B200_LIB=$PWD/dawnsearch_b200/$lib/libdawn_b200.so timeout 200 python tools/ab_gemm.py 12500000 1024 100 gemm_growth 0 2>&1 | tail -1
  done
done 2>&1 | tee $O/ab_arrive.txt
