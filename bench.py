#!/usr/bin/env python
"""bench.py -- exact top-k over a synthetic N x 384 fp16 corpus on 1..8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--batch B] [--k 10]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port) on the host cores

A step = one batch of B queries answered against the whole corpus (R rows, sharded by id range
over the N GPUs: strong scaling).  `value` = queries/s with queries and corpus resident in HBM
(CUDA events, max over ranks); `e2e` = the same through the C ABI with host buffers
(dawn_index_search_batch at N=1, ShardedIndex.search at N>1), H2D + D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xDA5EA2C4
DIM = 384
ROW_BYTES = 768


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="total corpus rows (all GPUs)")
    ap.add_argument("--batch", type=int, default=1, help="queries per step")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", default="", help="extra device-resident points: 'batch[:k],...' e.g. 1,4,16:100")
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def cpu_arm(args, steps, warmup, rows_total):
    """The reference arm / cpu_baseline: the oracle port (threaded SIMD exact scan, oracle/cpu_scan.c)
    on the host cores, on a bounded sample of the workload, extrapolated linearly (a scan is O(rows))."""
    from oracle import oracle as O

    sample = min(args.cpu_sample_rows, rows_total)
    stored = O.synth_rows_f16(SEED, 0, sample)
    qs = O.make_queries(SEED, SEED + 1, max(args.batch * (steps + warmup), 1), rows_total)
    threads = O.cpu_threads()
    times = []
    for s in range(warmup + steps):
        q = qs[s * args.batch:(s + 1) * args.batch]
        t0 = time.perf_counter()
        O.cpu_scan_f16(stored, None, q, args.k, threads=threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    scale = rows_total / sample
    step_s = statistics.mean(times) * scale
    qps = args.batch / step_s
    return {
        "value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
        "sample": f"{sample} of {rows_total} rows per step (oracle/cpu_scan.c, {threads} threads), "
                  f"time scaled x{scale:.1f} (scan is linear in rows); {steps} steps of batch {args.batch}",
        "ms_per_step": step_s * 1e3, "sample_ms_per_step": statistics.mean(times) * 1e3,
    }


def workload_name(args):
    return (f"exact top-{args.k} over {args.rows} x 384 fp16 (synthetic unit vectors), batch {args.batch} "
            f"per step, corpus sharded by id range over {args.gpus} GPU(s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_arm(args, steps, warmup, args.rows)
    line = {
        "impl": "reference", "metric": "queries/sec, exact top-k over N x 384 fp16", "value": r["value"],
        "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 accumulate over fp16 storage", "data": "synthetic",
        "config": {"workload": workload_name(args), "rows": args.rows, "batch": args.batch, "k": args.k,
                   "note": "reference's USearch 0.22.3 path cannot be built here (no cargo, crate not vendored); "
                           "this arm is the CPU exact-scan port on all host threads"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import dawnsearch_b200 as D
    from dawnsearch_b200 import synth as O  # workload generator (numpy twin of the device generator)
    from dawnsearch_b200.sharded import ShardedIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    shard = (args.rows + world - 1) // world
    first = rank * shard
    n_local = max(0, min(shard, args.rows - first))
    sh = ShardedIndex(local, max(n_local, 1))
    t0 = time.perf_counter()
    sh.index.add_synthetic(SEED, first, n_local)
    fill_s = time.perf_counter() - t0

    B, k, K, W = args.batch, args.k, args.steps, args.warmup
    qs_host = O.make_queries(SEED, SEED + 1, B * (K + W), args.rows, planted_fraction=0.5)
    qs_dev = torch.from_numpy(qs_host).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident queries, CUDA events on the launching stream -------------
    for s in range(W):
        sh.search_device(qs_dev[s * B:(s + 1) * B], k)
    barrier()
    sh.index.set_profiling(True)
    sh.index.profile(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record()
    for s in range(K):
        sh.search_device(qs_dev[(W + s) * B:(W + s + 1) * B], k)
        ev[s + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[K])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    prof = sh.index.profile(reset=True)
    sh.index.set_profiling(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * K / (total_ms / 1e3)

    # ---- e2e: host buffers through the public API, copies inside the timed region --------
    def e2e_call(q):
        if world == 1:
            return sh.index.search_batch(q, k)
        return sh.search(q, k)

    for s in range(min(W, 3)):
        e2e_call(qs_host[s * B:(s + 1) * B])
    barrier()
    lat = []
    t0 = time.perf_counter()
    for s in range(K):
        t1 = time.perf_counter()
        res = e2e_call(qs_host[(W + s) * B:(W + s + 1) * B])
        lat.append((time.perf_counter() - t1) * 1e3)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_qps = B * K / e2e_s
    prof_e2e = sh.index.profile(reset=True)

    # sanity: a planted neighbour must come back first (a wrong kernel cannot post a number)
    planted = O.planted_rows(SEED + 1, B * (K + W), args.rows)
    if len(planted):
        probe = e2e_call(qs_host[0:1])
        assert int(probe[0][0][0]) == int(planted[0]) + 1, "planted neighbour not returned first"
        sh.index.profile(reset=True)

    # ---- optional sweep over (batch, k): device-resident, 3 warm-up + 10 timed steps each ---
    sweep = []
    if args.sweep:
        for item in args.sweep.split(","):
            sb, sk = (int(x) for x in item.split(":")) if ":" in item else (int(item), k)
            q = torch.from_numpy(O.make_queries(SEED, SEED + 5, sb, args.rows)).to(dev)
            for _ in range(3):
                sh.search_device(q, sk)
            barrier()
            sh.index.set_profiling(True)
            sh.index.profile(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                sh.search_device(q, sk)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / 10
            p = sh.index.profile(reset=True)
            sh.index.set_profiling(False)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            pass_ms = p["scan_ms"] / max(int(p["scan_launches"]), 1)
            sweep.append({"batch": sb, "k": sk, "qps": sb / (ms / 1e3), "ms_per_step": ms,
                          "scan_passes": int(p["scan_launches"]) // 10, "scan_ms_per_pass": pass_ms,
                          "scan_gbps": n_local * ROW_BYTES / (pass_ms / 1e3) / 1e9 if pass_ms > 0 else None,
                          "finalize_ms": p["finalize_ms"] / max(int(p["finalize_launches"]), 1)})

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        scan_launches = max(int(prof["scan_launches"]), 1)
        scan_ms = prof["scan_ms"] / scan_launches
        algo_bytes = n_local * ROW_BYTES
        achieved = algo_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
        lat_sorted = sorted(lat)
        line = {
            "metric": "queries/sec, exact top-k over N x 384 fp16", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 accumulate over fp16 storage",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "rows": args.rows, "rows_per_gpu": n_local,
                       "batch": B, "k": k, "l2": "inputs larger than L2 (corpus shard >> 126 MB), no flush",
                       "corpus_fill_s": round(fill_s, 3)},
            "latency_ms": {"device_p50": statistics.median(step_ms), "device_max": max(step_ms),
                           "e2e_p50": statistics.median(lat), "e2e_p99": lat_sorted[min(len(lat) - 1, int(0.99 * len(lat)))]},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                    "d2h_bytes_per_step": B * k * 12 + B * 8 + 4, "ms_per_step": e2e_s / K * 1e3},
            "gpu_launches": int(prof["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "scan_topk_f16_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                         "avg_launch_ms": scan_ms, "launches_timed": int(prof["scan_launches"]),
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "finalize_avg_ms": prof["finalize_ms"] / max(int(prof["finalize_launches"]), 1)},
            "clocks": clocks,
            "exactness": {"uncertified_queries": int(prof["uncertified"]) + int(prof_e2e["uncertified"]),
                          "escalations": int(prof_e2e["escalations"])},
        }
        if sweep:
            line["batch_sweep"] = sweep
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_arm(args, steps=3, warmup=1, rows_total=args.rows)
            line["cpu_baseline"] = {kk: r[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
