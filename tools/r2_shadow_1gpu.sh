#!/bin/bash
# int8-shadow path on ONE B200: full GPU suite, ncu evidence (launch list of the default bench, full capture of the final round),
# memcheck, and the single-GPU lines of the BASELINE configs with the option on
set -u
O=gpurun_out/r2_shadow; mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log | cut -c1-200
(time timeout 600 python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err
timeout 300 python bench.py --rows 10000000 --no-cpu-baseline --sweep 1,2,4,1:20,1:100 > $O/bench_10m_c2.json 2> /dev/null
timeout 300 python bench.py --rows 12500000 --no-cpu-baseline --steps 20 --warmup 5 --sweep 1024:100,1024:10,1024:20,4096:10,256:10,1 > $O/bench_12m5.json 2> /dev/null
timeout 400 python bench.py --rows 50000000 --no-cpu-baseline --steps 10 --latency-steps 100 --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096 > $O/bench_50m_c4.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_default_bench_100m_b1024.csv \
    python bench.py --steps 2 --warmup 3 --latency-steps 10 --no-cpu-baseline --no-parity-check > $O/bench_under_ncu.log 2>&1
cap() {  # name regex skip kind rows batch k
  DAWN_OPTS=shadow_i8=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/$1 python tools/ncu_target.py $4 $5 $6 $7 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_ncu_full_raw.csv 2>/dev/null
}
cap gemm_i8_shadow_final_round_2cta gemm_i8_topk_kernel 11 f16gemm 20000000 1024 10
cap select_i8_shadow select_i8_kernel 12 f16gemm 20000000 1024 10
rm -f $O/*.ncu-rep
(echo "== memcheck shadow"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py shadow 2>&1 | tail -4) > $O/sanitizer_memcheck_shadow.txt 2>&1
cat $O/sanitizer_memcheck_shadow.txt
ls -la $O | head -30
