mkdir -p gpurun_out/r2p; O=gpurun_out/r2p
make -C dawnsearch_b200/csrc > $O/make.log 2>&1
for v in hint nohint; do
  if [ $v = nohint ]; then export DAWN_NO_CARVEOUT_HINT=1; else unset DAWN_NO_CARVEOUT_HINT; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_12m5_k10_$v.csv python tools/ncu_target.py f16gemm 12500000 1024 10 > $O/t_$v.log 2>&1
  # wall-clock per batch outside ncu (CUDA events around whole batches), 3 x 6 batches
  timeout 200 python tools/ab_gemm.py 12500000 1024 10 gemm_growth 0 2>&1 | tail -1
  timeout 200 python tools/ab_gemm.py 2000000 1024 10 gemm_growth 0 2>&1 | tail -1
done
