mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_f32.py -m gpu -x -q 2>&1 | tail -4) | tee $O/pytest.txt
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
for v in sync nosync; do
  if [ $v = sync ]; then export DAWN_OPTS=gemm_unit_sync=1; else export DAWN_OPTS=gemm_unit_sync=0; fi
  timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 9 -c 1 --csv --log-file $O/f16_$v.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16_$v.log 2>&1
  echo "== f16 $v"; grep -v "^==" $O/f16_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
  timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_i8_topk_kernel -s 11 -c 1 --csv --log-file $O/i8_$v.csv python tools/ncu_target.py i8gemm 20000000 1024 10 > $O/i8_$v.log 2>&1
  echo "== i8 $v"; grep -v "^==" $O/i8_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
done
unset DAWN_OPTS
# wall-clock A/B in one process, interleaved (thermal drift hits both)
timeout 300 python tools/ab_gemm.py 12500000 1024 10 gemm_unit_sync 1,0 2>&1 | tail -2
timeout 300 python tools/ab_gemm.py 100000000 1024 10 gemm_unit_sync 1,0 2>&1 | tail -2
DAWN_AB_SCALAR=i8 timeout 300 python tools/ab_gemm.py 62500000 1024 10 gemm_unit_sync 1,0 2>&1 | tail -2
