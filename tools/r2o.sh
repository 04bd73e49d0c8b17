mkdir -p gpurun_out/r2o; O=gpurun_out/r2o
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum
for b in 256 512 1024 2048; do
timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_i8_topk_kernel -s 11 -c 1 --csv --log-file $O/i8_b$b.csv python tools/ncu_target.py i8gemm 20000000 $b 10 > $O/i8_$b.log 2>&1
echo "== i8 batch $b"; grep -v "^==" $O/i8_b$b.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -7
done
timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 9 -c 1 --csv --log-file $O/f16_b1024.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16.log 2>&1
echo "== f16 batch 1024"; grep -v "^==" $O/f16_b1024.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tail -7
