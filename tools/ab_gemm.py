"""A/B of tensor-core-path tuning knobs on ONE GPU in ONE process (boxes differ by several percent)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dawnsearch_b200 as D
from dawnsearch_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
knob = sys.argv[4] if len(sys.argv) > 4 else "gemm_sequential_tiles"
values = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0, 1]
settings = [{knob: v} for v in values]
scalar = os.environ.get("DAWN_AB_SCALAR", "f16")
idx = D.new_index(D.IndexOptions(capacity=rows, quantization=D.ScalarKind.I8 if scalar == "i8" else D.ScalarKind.F16))
idx.add_synthetic(0xDA5EA2C4, 0, rows)
if os.environ.get("DAWN_AB_SHADOW", "0") == "1":
    idx.set_option("shadow_i8", 1)  # fp16 corpus filtered through its int8 copy
dev = torch.device("cuda:0")
q = torch.from_numpy(synth.make_queries(0xDA5EA2C4, 3, batch, rows)).to(dev)
L = torch.zeros((batch, k), dtype=torch.int64, device=dev); Dd = torch.zeros((batch, k), dtype=torch.float32, device=dev)
Cn = torch.zeros(batch, dtype=torch.int32, device=dev); Fl = torch.zeros(batch, dtype=torch.int32, device=dev)
def run(n):
    for _ in range(n):
        idx.search_device(q.data_ptr(), batch, k, L.data_ptr(), Dd.data_ptr(), Cn.data_ptr(), Fl.data_ptr(), 1)
    torch.cuda.synchronize()
run(3)
res = {}
for rep in range(3):
    for st in settings:
        for kk, v in st.items(): idx.set_option(kk, v)
        run(1)
        idx.set_profiling(True); idx.profile(reset=True)
        run(6)
        p = idx.profile(reset=True); idx.set_profiling(False)
        res.setdefault(json.dumps(st), []).append(p["gemm_ms"] / p["gemm_batches"])
for s, v in res.items():
    ms = sorted(v)[len(v) // 2]
    print(s, [round(x, 3) for x in v], "median", round(ms, 3), "TF", round(2 * batch * rows * 384 / ms / 1e9, 1))
