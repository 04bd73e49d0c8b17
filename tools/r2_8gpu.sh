#!/bin/bash
# Round-2 measurements on N B200 of one box (default 8): BASELINE configs C3 (k=100), C5 (500M int8), C4 (batch sweep at 50M),
# the headline config (k=10), and the one-process C-ABI front (dawn_multi_*, NCCL inside the library).
set -u
N=${1:-8}
O=gpurun_out/r2_${N}gpu
mkdir -p $O
make -C dawnsearch_b200/csrc > $O/make.log 2>&1; make -C oracle >> $O/make.log 2>&1
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29511 bench.py --gpus $N --k 100 --no-cpu-baseline > $O/c3_k100.json 2> $O/c3.err
timeout 400 $TR 29512 bench.py --gpus $N --scalar i8 --rows 500000000 --steps 10 --warmup 3 --latency-steps 100 --no-cpu-baseline \
    --sweep 1,2,16,256,1024:100,4096 > $O/c5_i8_500m.json 2> $O/c5.err
timeout 400 $TR 29513 bench.py --gpus $N --rows 50000000 --no-cpu-baseline --steps 10 --latency-steps 100 \
    --sweep 1,2,4,8,16,64,128,256,1024,4096 > $O/c4_sweep_50m.json 2> $O/c4.err
timeout 400 $TR 29514 bench.py --gpus $N --no-cpu-baseline > $O/headline_k10.json 2> $O/k10.err
timeout 400 python bench.py --front multi --gpus $N --steps 10 --warmup 3 --latency-steps 100 > $O/front_multi_k10.json 2> $O/multi.err
for f in c3 c5 c4 k10 multi; do echo "== $f"; tail -c 300 $O/$f.err; done
ls -la $O
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), "parity", (d.get("parity_check") or {}).get("bit_identical"), "roofline", round(d["roofline"]["frac"],3), (d.get("batch1") or {}).get("latency_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
