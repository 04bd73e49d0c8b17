mkdir -p gpurun_out/r2m; O=gpurun_out/r2m
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
for v in "base:" "chunk8:gemm_chunk_tiles=8" "chunk2:gemm_chunk_tiles=2" "seq:gemm_sequential_tiles=1" "cg1:gemm_cta_group=1"; do
  name=${v%%:*}; opts=${v#*:}
  DAWN_OPTS=$opts timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_i8_topk_kernel -s 11 -c 1 --csv --log-file $O/i8_$name.csv python tools/ncu_target.py i8gemm 20000000 1024 10 > $O/$name.log 2>&1
  echo "== $name"; grep -v "^==" $O/i8_$name.csv | cut -d, -f5,13- | tail -5
done
# the fp16 kernel for comparison
timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_topk_kernel -s 9 -c 1 --csv --log-file $O/f16_base.csv python tools/ncu_target.py f16gemm 20000000 1024 10 > $O/f16.log 2>&1
echo "== f16"; grep -v "^==" $O/f16_base.csv | cut -d, -f5,13- | tail -5
