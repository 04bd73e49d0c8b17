#!/bin/bash
# Round-2 ncu evidence on ONE B200: launch list of the default bench command, one full capture per dominant kernel
# (raw CSV exported next to it), compute-sanitizer memcheck over small shapes of every path.
set -u
O=gpurun_out/r2_prof
mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
# 1. every launch of the default bench workload with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default_bench_100m_b1024.csv \
    python bench.py --steps 2 --warmup 3 --latency-steps 10 --no-cpu-baseline --no-parity-check > $O/bench_under_ncu.log 2>&1
# 2. full captures
cap() {  # name regex skip kind rows batch k
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/$1 python tools/ncu_target.py $4 $5 $6 $7 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_ncu_full_raw.csv 2>/dev/null
}
cap gemm_f16_final_round_2cta gemm_topk_kernel 9 f16gemm 20000000 1024 10
cap gemm_i8_final_round_2cta gemm_i8_topk_kernel 11 i8gemm 20000000 1024 10
cap select_i8 select_i8_kernel 12 i8gemm 20000000 1024 10
cap scan_f16_q1_10m scan_topk_f16 1 f16scan 10000000 1 10
cap scan_i8_q1_20m scan_topk_i8 1 i8scan 20000000 1 10
rm -f $O/*.ncu-rep
# 3. sanitizer
for w in scan gemm i8 i8gemm f32 multi; do
  echo "== memcheck $w"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py $w 2>&1 | tail -4
done > $O/sanitizer_memcheck.txt 2>&1
tail -30 $O/sanitizer_memcheck.txt
ls -la $O
