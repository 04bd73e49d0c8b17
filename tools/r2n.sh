mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_f32.py -m gpu -x -q 2>&1 | tail -6) | tee $O/pytest.txt
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_i8_topk_kernel -s 11 -c 1 --csv --log-file $O/i8_sched.csv python tools/ncu_target.py i8gemm 20000000 1024 10 > $O/i8.log 2>&1
echo "== i8"; grep -v "^==" $O/i8_sched.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 7 -c 1 --csv --log-file $O/f16_sched.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16.log 2>&1
echo "== f16"; grep -v "^==" $O/f16_sched.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
timeout 200 python tools/i8_tensor_bench.py 62500000 1024 1 > $O/i8_bench.json 2>/dev/null; cat $O/i8_bench.json; echo
timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 0 2>&1 | tail -1
timeout 200 python tools/ab_gemm.py 12500000 1024 100 gemm_growth 0 2>&1 | tail -1
timeout 200 python tools/ab_gemm.py 100000000 1024 10 gemm_growth 0 2>&1 | tail -1
