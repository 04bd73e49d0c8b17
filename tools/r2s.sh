mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4) | tee $O/pytest.txt
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
for v in slots noslots; do
  if [ $v = noslots ]; then export DAWN_NO_DIE_SLOTS=1; else unset DAWN_NO_DIE_SLOTS; fi
  DAWN_DEBUG_DIE_SLOTS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_i8_topk_kernel -s 11 -c 1 --csv --log-file $O/i8_$v.csv python tools/ncu_target.py i8gemm 20000000 1024 10 > $O/i8_$v.log 2>&1
  grep "dawn\]" $O/i8_$v.log | cut -c1-240
  echo "== i8 $v"; grep -v "^==" $O/i8_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
  timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 9 -c 1 --csv --log-file $O/f16_$v.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16_$v.log 2>&1
  echo "== f16 $v"; grep -v "^==" $O/f16_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5
done
