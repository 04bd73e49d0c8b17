mkdir -p gpurun_out/r2t; O=gpurun_out/r2t
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
for v in sync nosync; do
  if [ $v = sync ]; then export DAWN_UNIT_SYNC=1; else unset DAWN_UNIT_SYNC; fi
  timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 9 -c 1 --csv --log-file $O/f16_$v.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16_$v.log 2>&1
  echo "== f16 $v"; grep -v "^==" $O/f16_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -5; tail -1 $O/f16_$v.log | cut -c1-150
  timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 0 2>&1 | tail -1
done
