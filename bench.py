#!/usr/bin/env python
"""bench.py -- exact top-k over a synthetic N x 384 fp16 corpus on 1..8 B200 (BASELINE.json metric:
queries/sec and p50 latency, exact top-10 over 100M x 384 fp16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--batch B] [--k 10]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port) on the host cores

A step = one batch of B queries (default 1024: the throughput configuration) answered against the
whole corpus (R rows, sharded by id range over the N GPUs: strong scaling).
  value  = queries/s with queries and corpus resident in HBM (CUDA events on the launching
           stream, max over ranks)
  e2e    = the same through the public API with HOST buffers (dawn_index_search_batch at N=1,
           ShardedIndex.search at N>1): H2D of the queries and D2H of the results inside the
           timed region
  batch1 = the latency configuration (one query per step): device and end-to-end p50/p99 and the
           HBM roofline of the streaming-scan kernel
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
METRIC = "queries/sec, exact top-k over N x 384 fp16"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="total corpus rows (all GPUs)")
    ap.add_argument("--batch", type=int, default=1024, help="queries per step")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--scalar", default="f16", choices=["f16", "i8"], help="stored precision of the corpus")
    ap.add_argument("--latency-steps", type=int, default=200, help="batch-1 steps for the latency section (0 = skip)")
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", default="", help="extra device-resident points: 'batch[:k],...' e.g. 1,4,16:100")
    return ap.parse_args()


def ncu_traffic_ratio(kernel_key: str):
    """DRAM bytes / algorithmic bytes of a kernel from the committed ncu captures (profiles/ncu_summary.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return float(json.load(f)[kernel_key]["traffic_over_algorithmic"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PF burst, ~1.4 PF sustained)"}


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
    on the host cores, on a bounded sample of the workload, scaled linearly (a scan is O(rows))."""
    from oracle import oracle as O

    sample = min(args.cpu_sample_rows, rows_total)
    stored = O.synth_rows_f16(SEED, 0, sample)
    qs = O.make_queries(SEED, SEED + 1, args.batch, rows_total)
    threads = O.cpu_threads()
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.cpu_scan_f16(stored, None, qs, args.k, threads=threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    scale = rows_total / sample
    step_s = statistics.mean(times) * scale
    return {
        "value": args.batch / step_s, "unit": "queries/s", "cores": threads, "kind": "port",
        "sample": f"{sample} of {rows_total} rows per step (oracle/cpu_scan.c, {threads} threads), "
                  f"time scaled x{scale:.1f} (a scan is linear in rows); {steps} steps of batch {args.batch}",
        "ms_per_step": step_s * 1e3, "sample_ms_per_step": statistics.mean(times) * 1e3,
    }


def workload_name(args):
    return (f"exact top-{args.k} over {args.rows} x 384 {'fp16' if args.scalar == 'f16' else 'int8 (+f32 scale per vector)'} "
            f"(synthetic unit vectors), batch {args.batch} "
            f"per step, corpus sharded by id range over {args.gpus} GPU(s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 1))
    r = cpu_arm(args, steps, warmup, args.rows)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"],
        "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 accumulate over fp16 storage", "data": "synthetic",
        "config": {"workload": workload_name(args), "rows": args.rows, "batch": args.batch, "k": args.k,
                   "note": "reference's USearch 0.22.3 path cannot be built here (no cargo, crate not vendored); "
                           "this arm is the CPU exact-scan port on all host threads"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print_json(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from dawnsearch_b200 import synth as O  # workload generator (numpy twin of the device generator)
    from dawnsearch_b200.sharded import ShardedIndex, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    first, n_local = shard_range(rank, world, args.rows)
    row_bytes = ROW_BYTES if args.scalar == "f16" else 388  # i8: 384 B + f32 scale
    sh = ShardedIndex(local, max(n_local, 1), quantization=0 if args.scalar == "f16" else 1)
    t0 = time.perf_counter()
    sh.index.add_synthetic(SEED, first, n_local)
    fill_s = time.perf_counter() - t0

    B, k, K, W = args.batch, args.k, args.steps, args.warmup
    # a pool of distinct query batches, cycled (fresh queries every step)
    n_pool = min(K + W, 8)
    pool_host = [O.make_queries(SEED, SEED + 1 + i, B, args.rows, planted_fraction=0.25) for i in range(n_pool)]
    pool_dev = [torch.from_numpy(q).to(dev) for q in pool_host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(queries_dev, kk, steps, warm, sample_clocks=False):
        """Device-resident timing: CUDA events on the launching stream + the library's own per-kernel events."""
        # back-to-back batches: the exchange of batch i overlaps the search of batch i+1 (not worth it
        # for single queries, where the side-stream bookkeeping costs more than the 12 us exchange)
        pipe = world > 1 and queries_dev[0].shape[0] >= 16
        for s in range(warm):
            sh.search_device(queries_dev[s % len(queries_dev)], kk, pipelined=pipe)
        sh.wait_results()
        barrier()
        sh.index.set_profiling(True)
        sh.index.profile(reset=True)
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for s in range(steps):
            sh.search_device(queries_dev[(warm + s) % len(queries_dev)], kk, pipelined=pipe)
            if s == steps - 1:
                sh.wait_results()  # the last event must cover the last batch's exchange + merge
            ev[s + 1].record()
        barrier()
        clocks = sampler.stop() if sampler else None
        total_ms = allmax(ev[0].elapsed_time(ev[steps]))
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        prof = sh.index.profile(reset=True)
        sh.index.set_profiling(False)
        return total_ms, step_ms, prof, clocks

    def e2e_call(q, kk):
        if world == 1:
            return sh.index.search_batch(q, kk)
        return sh.search(q, kk)

    def timed_e2e(queries_host, kk, steps, warm):
        for s in range(warm):
            e2e_call(queries_host[s % len(queries_host)], kk)
        barrier()
        lat = []
        t0 = time.perf_counter()
        for s in range(steps):
            t1 = time.perf_counter()
            e2e_call(queries_host[(warm + s) % len(queries_host)], kk)
            lat.append((time.perf_counter() - t1) * 1e3)
        barrier()
        return allmax(time.perf_counter() - t0), lat

    def roofline_of(prof, batch, peaks):
        """Roofline of the dominant kernel of a device-resident run."""
        if prof["gemm_batches"]:
            gms = prof["gemm_ms"] / int(prof["gemm_batches"])
            flops = 2.0 * batch * n_local * DIM
            ach = flops / (gms / 1e3) / 1e12
            ratio = ncu_traffic_ratio("gemm_topk_kernel<2> final round")
            return {"bound": "tensor", "kernel": "gemm_topk_kernel (tcgen05; all rounds of a batch incl. select)",
                    "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tf_sustained"],
                    "traffic": ratio * n_local * row_bytes if ratio else None,
                    "traffic_note": "DRAM bytes per batch = ncu dram bytes / corpus bytes of the captured round "
                                    "(profiles/ncu_summary.json) x this shard's corpus bytes",
                    "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "frac_of_burst_peak": ach / peaks["tf_burst"], "frac_of_nominal_2250": ach / 2250.0,
                    "algorithmic_flops_per_batch": flops, "avg_batch_ms": gms,
                    "corpus_stream_gbps": n_local * row_bytes / (gms / 1e3) / 1e9}
        launches = max(int(prof["scan_launches"]), 1)
        sms = prof["scan_ms"] / launches
        algo = n_local * row_bytes
        ach = algo / (sms / 1e3) / 1e9 if sms > 0 else 0.0
        ratio = ncu_traffic_ratio("scan_topk_f16_kernel<1>" if args.scalar == "f16" else "scan_topk_i8_kernel<1>")
        return {"bound": "hbm", "kernel": "scan_topk_f16_kernel" if args.scalar == "f16" else "scan_topk_i8_kernel", "achieved": ach, "peak": peaks["hbm"],
                "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": ratio * algo if ratio else None,
                "traffic_note": "ncu dram bytes / algorithmic bytes of the captured launch (profiles/ncu_summary.json) x this launch's algorithmic bytes",
                "peak_source": peaks["source"] + " hbm_gbs", "frac_of_nominal_8TBs": ach / 8000.0,
                "algorithmic_bytes_per_launch": algo, "avg_launch_ms": sms, "launches_timed": int(prof["scan_launches"])}

    peaks = measured_peaks()

    # ---- latency configuration: one query per step -----------------------------------------
    # (measured BEFORE the throughput run: that one leaves the chip on its 1000 W power cap for seconds,
    #  and single-query latency is a separate workload, not a tail of it)
    batch1 = None
    if args.latency_steps > 0 and B != 1:
        q1_host = [pool_host[0][i:i + 1] for i in range(min(B, 16))]
        q1_dev = [pool_dev[0][i:i + 1] for i in range(min(B, 16))]
        t_ms, s_ms, p1, _ = timed_device(q1_dev, k, args.latency_steps, 20)
        e_s, l1 = timed_e2e(q1_host, k, args.latency_steps, 20)
        l1s = sorted(l1)
        batch1 = {"qps_device": args.latency_steps / (t_ms / 1e3), "qps_e2e": args.latency_steps / e_s,
                  "latency_ms": {"device_p50": statistics.median(s_ms), "device_max": max(s_ms),
                                 "e2e_p50": statistics.median(l1), "e2e_p99": l1s[min(len(l1s) - 1, int(0.99 * len(l1s)))]},
                  "roofline": roofline_of(p1, 1, peaks)}

    # ---- main workload -------------------------------------------------------------------
    total_ms, step_ms, prof, clocks = timed_device(pool_dev, k, K, W, sample_clocks=True)
    value = B * K / (total_ms / 1e3)
    e2e_s, lat = timed_e2e(pool_host, k, K, min(W, 3))
    prof_e2e = sh.index.profile(reset=True)

    # sanity: a planted neighbour must come back first (a wrong kernel cannot post a number)
    planted = O.planted_rows(SEED + 1, B, args.rows, 0.25)
    if len(planted):
        probe = e2e_call(pool_host[0], k)
        assert int(probe[0][0][0]) == int(planted[0]) + 1, "planted neighbour not returned first"
        sh.index.profile(reset=True)

    # ---- optional sweep over (batch, k): device-resident, 3 warm-up + 10 timed steps each ---
    sweep = []
    if args.sweep:
        for item in args.sweep.split(","):
            sb, sk = (int(x) for x in item.split(":")) if ":" in item else (int(item), k)
            q = [torch.from_numpy(O.make_queries(SEED, SEED + 5, sb, args.rows)).to(dev)]
            t_ms, _, p, _ = timed_device(q, sk, 10, 3)
            ms = t_ms / 10
            ent = {"batch": sb, "k": sk, "qps": sb / (ms / 1e3), "ms_per_step": ms,
                   "finalize_ms": p["finalize_ms"] / max(int(p["finalize_launches"]), 1)}
            r = roofline_of(p, sb, peaks)
            if p["gemm_batches"]:
                ent.update({"path": "tcgen05", "gemm_ms": r["avg_batch_ms"], "tflops": r["achieved"],
                            "corpus_gbps": r["corpus_stream_gbps"]})
            else:
                ent.update({"path": "scan", "scan_passes": int(p["scan_launches"]) // 10,
                            "scan_ms_per_pass": r["avg_launch_ms"], "scan_gbps": r["achieved"]})
            sweep.append(ent)

    if rank == 0:
        lat_sorted = sorted(lat)
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16 x f16 -> f32 (tensor cores) / f32 scan, exact f32 re-score",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "rows": args.rows, "rows_per_gpu": n_local,
                       "batch": B, "k": k, "l2": "inputs larger than L2 (corpus shard >> 126 MB), no flush",
                       "multi_gpu": ("device-resident `value`: the all-gather + merge of batch i run on a side stream "
                                     "while batch i+1 is searched; e2e: synchronous per call") if world > 1 else None,
                       "corpus_fill_s": round(fill_s, 3)},
            "latency_ms": {"device_step_p50": statistics.median(step_ms), "device_step_max": max(step_ms),
                           "e2e_step_p50": statistics.median(lat),
                           "e2e_step_p99": lat_sorted[min(len(lat) - 1, int(0.99 * len(lat)))]},
            "e2e": {"value": B * K / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                    "d2h_bytes_per_step": B * k * 12 + B * 8 + 4, "ms_per_step": e2e_s / K * 1e3},
            "gpu_launches": int(prof["kernel_launches"]),
            "roofline": roofline_of(prof, B, peaks),
            "clocks": clocks,
            "exactness": {"uncertified_queries": int(prof_e2e["uncertified"]),
                          "escalated_to_exact_scan": int(prof_e2e["escalations"]),
                          "note": "labels and distances bit-identical to the CPU oracle (tests/); the e2e path "
                                  "re-runs any query whose exactness certificate fails through the f32 scan"},
        }
        try:  # recall of the reference's kind of index vs the exact answer (config C1), measured by
            # tools/c1_reference_path.py on a B200 box and committed; not re-measured here (52 s build)
            with open(os.path.join(ROOT, "profiles", "r01_c1_reference_path_hnsw_recall.json")) as f:
                c1 = json.load(f)
            line["reference_path_recall"] = {
                "source": "profiles/r01_c1_reference_path_hnsw_recall.json (tools/c1_reference_path.py)",
                "index": c1["index"], "config": c1["config"],
                "hnsw_recall_at_10_ef64": c1["results"]["k10"]["hnsw_ef64"]["recall"],
                "hnsw_p50_us_ef64": c1["results"]["k10"]["hnsw_ef64"]["p50_us"],
                "exact_gpu_recall": 1.0, "exact_gpu_p50_us": c1["b200_exact_c_abi"]["p50_us"], "note": c1["note"]}
        except Exception:
            pass
        if batch1:
            line["batch1"] = batch1
        if sweep:
            line["batch_sweep"] = sweep
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_arm(args, steps=2, warmup=1, rows_total=args.rows)
            line["cpu_baseline"] = {kk: r[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        print_json(line)
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    # Only the JSON line may reach stdout (NCCL and torchrun print banners there): park the real
    # stdout, point fd 1 at stderr while working, and restore it for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    line_holder = []
    global print_json
    def print_json(obj):
        line_holder.append(json.dumps(obj))
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in line_holder:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
