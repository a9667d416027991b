#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Vettore scan path.

Metric (BASELINE.json): queries/sec @k=10, 1M x 768 fp32 cosine flat exact scan.
A "step" is one single-query search over the resident corpus (configs[1], batch of 1 —
the HBM-bound case the north_star's >=80%-of-roofline target is stated on).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # CPU restatement of the reference

N>1 is launched by the driver under torchrun (one rank per GPU). The corpus is row-sharded:
every rank holds its own 1M x 768 shard (weak scaling: per-GPU work fixed, corpus = N M rows),
the query is replicated, per-rank top-k records are all-gathered over NVLink (NCCL) and
merged by the K7 kernel. `value` = shard scans completed by all ranks per second
(= queries/s x N; at N=1 exactly BASELINE's queries/s).

Only the `cpu_baseline` leg and `--impl reference` touch oracle/ (the CPU checker).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec @k=10, 1M x 768 fp32 cosine flat"
SEED = 20_260_721  # the reference's own bench seed (bench/search_modes_bench.exs:14)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="rows per GPU shard")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=64, help="distinct queries rotated through the steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons (NVML, every 20 ms) while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop.is_set():
                self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                     pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                self.stop.wait(0.02)
        except Exception:
            self._run_smi()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6 and p[0].isdigit():
                    self.max_mhz = int(p[1])
                    self.samples.append((int(p[0]), sum(b for b, v in zip(bits, p[2:6]) if v == "Active")))
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        time.sleep(0.05)           # let NVML initialise before the timed region starts
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        mhz = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[1]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if mask & b], "samples": len(self.samples)}


def make_rows_torch(rows: int, dim: int, seed: int, device):
    """i.i.d. standard normal rows, L2-normalised the reference's way (f64 norm, then f32)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn(rows, dim, generator=g, device=device, dtype=torch.float32)
    chunk = 131072
    for s in range(0, rows, chunk):  # distances.rs:350-361: divide in f64, cast to f32
        blk = x[s:s + chunk].double()
        x[s:s + chunk] = (blk / blk.norm(dim=1, keepdim=True)).float()
    return x


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: the Rust NIF cannot
    be built in this image), all host threads, each step = `threads` concurrent single-query
    scans (one sequential scan per dirty-scheduler call, nifs.rs:297-309)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch

    import oracle

    threads = os.cpu_count() or 1
    torch.manual_seed(SEED)
    rows = make_rows_torch(args.rows, args.dim, SEED, torch.device("cpu")).numpy()
    queries = make_rows_torch(max(threads, 8), args.dim, SEED + 1, torch.device("cpu")).numpy()
    steps = max(1, min(args.steps, 8))      # bounded: each step scans the corpus `threads` times
    warmup = max(1, min(args.warmup, 1))
    for _ in range(warmup):
        oracle.flat_scan_timed("cosine", rows, queries[:threads], args.k, threads)
    t = 0.0
    for s in range(steps):
        q = np.roll(queries, s, axis=0)[:threads]
        secs, _ = oracle.flat_scan_timed("cosine", rows, q, args.k, threads)
        t += secs
    qps = steps * threads / t
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"flat cosine exact scan {args.rows}x{args.dim} fp32, single query, k={args.k}",
                   "note": "CPU restatement of flat.rs:96-124 (oracle port); one step = one query per host thread"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x {threads} concurrent single-query scans of the full "
                                   f"{args.rows}x{args.dim} corpus"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- CUDA arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from vettore_b200 import nifs
    from vettore_b200.sharded import ShardedFlat, set_global_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, d, k = args.rows, args.dim, args.k
    # ---- corpus shard: generated on the device, brought to the host once, then ingested
    # through the reference-facing boundary (flat_insert_many -> bulk H2D).
    x = make_rows_torch(n, d, SEED + 7919 * rank, dev)
    host_rows = x.cpu().numpy()
    del x
    torch.cuda.empty_cache()
    base = rank * n
    ids = [f"{base + i:09d}" for i in range(n)]
    index = nifs.flat_new_cosine()
    t0 = time.perf_counter()
    res = nifs.flat_insert_matrix(index, ids, host_rows)
    assert res == ("ok", ()), res
    ingest_s = time.perf_counter() - t0
    if world > 1:
        set_global_ranks(index, base, n)
    queries = make_rows_torch(args.queries, d, SEED + 1, dev)          # identical on every rank
    q_host = queries.cpu().pin_memory()
    sharded = ShardedFlat(index, k=k, nq=1, device=dev)

    # ---- parity spot check against the oracle before timing (rank 0, its own shard)
    if rank == 0:
        import oracle
        st, hits = nifs.flat_search(index, q_host[0].numpy(), k)
        assert st == "ok", hits
        sub = 100_000
        st2, ref = oracle.flat_search_dense("cosine", host_rows[:sub], ids[:sub], q_host[0].numpy(), k)
        ref_ids = {h[0] for h in ref}
        got_sub = [h for h in hits if int(h[0]) - base < sub]
        assert all(h[0] in ref_ids for h in got_sub), "parity spot check failed"

    # ---- value: device-timed, query already resident in HBM
    def step_device(i):
        return sharded.search_device(queries[i % args.queries: i % args.queries + 1])

    for i in range(args.warmup):
        step_device(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for i in range(args.steps):
            step_device(i)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t_dev = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total = float(t_dev.item())

    # ---- scan-kernel-only timing (roofline): the local scan without the exchange
    import ctypes as C
    from vettore_b200._lib import lib
    lay = sharded.layout
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def scan_only(i):
        q = queries[i % args.queries: i % args.queries + 1]
        rc = lib().vb_flat_search_device(index.handle, C.c_void_p(q.data_ptr()), 1, d, k,
                                         C.c_void_p(sharded.local.data_ptr() + lay["keys"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["values"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["rows"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["counts"]), stream)
        assert rc == 0

    barrier()
    ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek0.record()
    for i in range(args.steps):
        scan_only(i)
    ek1.record()
    torch.cuda.synchronize()
    kernel_ms = ek0.elapsed_time(ek1) / args.steps

    # ---- e2e: host query in, host result out, through the public host-facing call
    def step_e2e(i):
        qi = i % args.queries
        if world == 1:
            st, hits = nifs.flat_search(index, q_host[qi].numpy(), k)   # C ABI: H2D + scan + D2H + ids
            return hits
        return sharded.search(q_host[qi:qi + 1])

    for i in range(max(3, args.warmup // 4)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())

    # ---- configs[1] also names a batch of 1024 queries: K2 (tcgen05 3xTF32 GEMM + fused top-k),
    # device-timed with the queries resident in HBM; reported inside `config`, not as the headline.
    batch = None
    if world == 1:
        nqb = 1024
        bq = make_rows_torch(nqb, d, SEED + 2, dev)
        bkeys = torch.zeros(nqb * k, dtype=torch.int64, device=dev)
        bvals = torch.zeros(nqb * k, dtype=torch.float32, device=dev)
        brows = torch.zeros(nqb * k, dtype=torch.int32, device=dev)
        bcnts = torch.zeros(nqb, dtype=torch.int32, device=dev)

        def batch_step():
            rc = lib().vb_flat_search_device(index.handle, C.c_void_p(bq.data_ptr()), nqb, d, k,
                                             C.c_void_p(bkeys.data_ptr()), C.c_void_p(bvals.data_ptr()),
                                             C.c_void_p(brows.data_ptr()), C.c_void_p(bcnts.data_ptr()), stream)
            assert rc == 0

        for _ in range(2):
            batch_step()
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(5):
            batch_step()
        b1.record()
        torch.cuda.synchronize()
        bms = b0.elapsed_time(b1) / 5
        # spot parity: query 0 of the batch through the single-query kernel
        st, one = nifs.flat_search(index, bq[0].cpu().numpy(), k)
        got_rows = brows[:k].cpu().numpy().tolist()
        assert [int(h[0]) - base for h in one] == got_rows, "batched and single-query results differ"
        batch = {"queries": nqb, "ms_per_batch": bms, "queries_per_sec": nqb / (bms * 1e-3),
                 "tf32_tflops_issued": 3 * 2.0 * nqb * n * d / (bms * 1e-3) / 1e12,
                 "kernel": "vb::flat_gemm_topk_kernel (tcgen05 3xTF32) + exact re-scoring"}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = n * d * 4
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC,
            "value": world * args.steps / (ms_total * 1e-3),
            "unit": "queries/s" if world == 1 else "1M-row shard scans/s (queries/s x n_gpus)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"flat cosine exact scan {n}x{d} fp32 per GPU, batch of 1 query, k={k}",
                "corpus_rows_total": n * world, "parallelism": f"row-shard x{world}",
                "l2_policy": f"corpus shard {alg_bytes / 1e9:.2f} GB >> 126 MB L2, {args.queries} rotating queries",
                "queries_per_sec": args.steps / (ms_total * 1e-3),
                "ingest_seconds_per_shard": round(ingest_s, 3),
                "batch_1024": batch,
            },
            "clocks": clocks.summary(),
            "e2e": {"value": world * args.steps / e2e_s, "unit": "queries/s" if world == 1 else "shard scans/s",
                    "h2d_bytes_per_step": d * 4, "d2h_bytes_per_step": k * 8 + 8 if world == 1 else int(sharded.out.numel()),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": args.steps * sharded.launches_per_search,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed
                         # ncu --set full capture (profiles/r1_flat_stream_ncu_summary.txt); default shape only
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                         # (profiles/r1_flat_stream_ncu_summary.txt: 3.073186 GB + 6.35 MB)
                         "traffic": 3.0795e9 if (n, d, k) == (1_000_000, 768, 10) else None,
                         "peak_source": peak_src,
                         "kernel": "vb::flat_stream_kernel<cosine, NV=6, RPW=1, W=16> (TMA-staged ring; the timed "
                                   "launch pair also holds the ~3 us unpack kernel)",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            threads = os.cpu_count() or 1
            nq = 4 * threads
            secs, _ = oracle.flat_scan_timed("cosine", host_rows, q_host[:nq].numpy() if nq <= args.queries
                                             else np.tile(q_host.numpy(), (nq // args.queries + 1, 1))[:nq], k, threads)
            line["cpu_baseline"] = {"value": nq / secs, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": f"{nq} single-query scans of the same {n}x{d} corpus, {threads} host "
                                              f"threads, one sequential scan per query (flat.rs:96-124 restated)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
