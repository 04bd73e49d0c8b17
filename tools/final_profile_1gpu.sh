#!/bin/bash
# Round-end ncu evidence + the default bench line on ONE B200.
set -u
O=gpurun_out/final
mkdir -p $O
timeout 600 python bench.py > $O/bench_100m_1gpu.json 2> $O/bench_100m_1gpu.err
timeout 200 python tools/ingest_rate.py > $O/ingest_rate_2m.json 2> $O/ingest_rate.err
B1="python bench.py --rows 10000000 --batch 1 --steps 40 --warmup 5 --latency-steps 0 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_topk_f16 --launch-skip 30 -c 1 -f -o $O/prof_scan_q1_10m $B1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:finalize_kernel --launch-skip 30 -c 1 -f -o $O/prof_finalize_scanpath $B1 > /dev/null 2>&1
ncu -i $O/prof_scan_q1_10m.ncu-rep --page raw --csv > $O/scan_q1_10m_ncu_full_raw.csv 2>/dev/null
ncu -i $O/prof_finalize_scanpath.ncu-rep --page raw --csv > $O/finalize_scanpath_ncu_full_raw.csv 2>/dev/null
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_bench.csv python bench.py --steps 2 --warmup 1 --latency-steps 5 --no-cpu-baseline > /dev/null 2>&1
rm -f $O/*.ncu-rep
ls -la $O | tail -12
