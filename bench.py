#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Vettore scan path.

Metric (BASELINE.json): queries/sec @k=10, 1M x 768 fp32 cosine flat exact scan.
A "step" is one single-query search over the resident corpus (configs[1], batch of 1 —
the HBM-bound case the north_star's >=80%-of-roofline target is stated on).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # CPU restatement of the reference

N>1 is launched by the driver under torchrun (one rank per GPU). The headline is weak scaling: every
rank holds its own 1M x 768 shard (corpus = N M rows), the query is replicated, the per-rank top-k
records are exchanged over NVLink (peer-memory stores, or one NCCL all-gather) and merged by the K7
kernel. `value` = 1M-row shard scans completed by all ranks per second (= queries/s over the N M-row
corpus x N; at N=1 exactly BASELINE's queries/s); the unit string stays "queries/s" so the driver can
set it against the reference arm (which scans the same 1M x 768 shard on the host cores).

Next to the headline every line carries, in `config`, the other BASELINE.json configurations at their
stated sizes, sharded over the N GPUs:
  c2_batch_1024  1M x 768 cosine, 1024-query batch, k=10 (K2, tcgen05 3xTF32)          [N=1 only]
  c3  flat inner product, 100M x 768 over N GPUs, 1024-query batch, k=100 (K2 + all-gather + K7)
  c4  quantized_search: 100M x 1024-bit sign codes, 1000 candidates -> exact cosine rerank to k=10
  c5  ColBERT MaxSim: 1M docs x 128 tokens x 128 dims, 32-token query, k=10
each with device-timed ms per step, queries/s, per-GPU GB/s or TFLOP/s, roofline fraction, the
rows/docs actually resident per GPU (and whether that is the full stated corpus), and the speed-up over
one GPU working through the same corpus shard by shard (N x local scan time / sharded step time).

Only the parity checks, the `cpu_baseline` legs and `--impl reference` touch oracle/ (the CPU checker).
"""
from __future__ import annotations

import argparse
import ctypes as C
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
TOL = 1e-5         # north_star: float scores within 1e-5 relative
INGEST_CHUNK = 1_000_000   # rows generated per device chunk (one seed per chunk)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="rows per GPU shard (headline)")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=64, help="distinct queries rotated through the steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="c2,c3,c4,c5", help="comma list of the extra BASELINE configs to run ('' = none)")
    ap.add_argument("--corpus-scale", type=float, default=1.0,
                    help="scales the 100M / 1M-doc corpora of c3/c4/c5 (development runs only; 1.0 = BASELINE sizes)")
    return ap.parse_args()


def measured_tf32_peak():
    """Dense TF32 tensor-pipe peak of this device, measured live with a bare tcgen05.mma kind::tf32 loop
    (vb_debug_tf32_peak, csrc/tc_probe.cu): the denominator SURVEY.md §8(d) asks for. None when it cannot run."""
    try:
        from vettore_b200._lib import lib
        fn = lib().vb_debug_tf32_peak
        fn.restype = C.c_int
        out = C.c_float(0.0)
        rc = fn(C.c_int(20000), C.byref(out))
        return float(out.value) if rc == 0 and out.value > 0 else None
    except Exception:
        return None


def workload_string(rows, dim, k):
    return f"flat cosine exact scan {rows}x{dim} fp32 per GPU, batch of 1 query, k={k}"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md 6.65 TB/s)", "bf16": 1655.0, "bf16_sustained": 1373.0,
           "tc_src": "fallback"}
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            out.update(hbm=float(j["hbm_gbs"]), hbm_src="measured (MEASURED_PEAKS.json hbm_gbs)")
            out.update(bf16=float(j["bf16_tflops"]), bf16_sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                       tc_src="measured (MEASURED_PEAKS.json bf16_tflops / 2 = dense TF32)")
        except Exception:
            pass
    return out


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
        capped = sum(1 for s in self.samples if s[1] & 0x4)
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_min_mhz": mhz[0], "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if mask & b], "samples": len(self.samples),
                "power_cap_sample_frac": round(capped / len(self.samples), 3)}


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


def close(a: float, b: float, tol: float = TOL) -> bool:
    """distances.rs:487-493 assert_close: |a-b| <= tol * max(1, |a|, |b|)."""
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


def assert_hits_match(actual, expected, what: str):
    """[(row or id, value)] lists: values within TOL, order equal except ties inside the tolerance and
    swaps across the cut with a value tying the last expected one (tests/helpers.py, north_star's bar)."""
    assert len(actual) == len(expected), (what, len(actual), len(expected))
    exp = dict(expected)
    last = expected[-1][1] if expected else 0.0
    for pos, ((ia, va), (ie, ve)) in enumerate(zip(actual, expected)):
        assert close(va, ve), (what, pos, ia, va, ie, ve)
        if ia != ie and ia not in exp:
            assert close(va, last), (what, pos, ia, va, "not in the oracle list and not a boundary tie", last)


def subsample_check(got, oracle_hits, sample_rows: int, descending: bool, what: str):
    """Size-independent parity property for a corpus too large for the CPU oracle: `got` = the CUDA path's
    [(global row, value)] over the WHOLE corpus, `oracle_hits` = the oracle's top-k over the first
    `sample_rows` rows only. (1) the list is sorted; (2) every CUDA hit that falls inside the sample carries
    the oracle's value for that row; (3) every oracle hit that beats the CUDA list's last value by more than
    the tolerance is present in the CUDA list. Returns the number of hits verified by value."""
    sgn = -1.0 if descending else 1.0
    vals = [sgn * v for _, v in got]
    assert all(vals[i] <= vals[i + 1] + TOL * max(1.0, abs(vals[i])) for i in range(len(vals) - 1)), (what, "unsorted")
    ref = dict(oracle_hits)
    got_rows = {r for r, _ in got}
    worst = vals[-1] if vals else float("inf")
    checked = 0
    for r, v in got:
        if r < sample_rows and r in ref:
            assert close(v, ref[r]), (what, r, v, ref[r])
            checked += 1
    for r, v in oracle_hits:
        if sgn * v < worst - TOL * max(1.0, abs(worst)):
            assert r in got_rows, (what, "oracle hit missing from the CUDA list", r, v, worst)
    return checked


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: the Rust NIF cannot be built in this
    image), all host threads, each step = `threads` concurrent single-query scans (one sequential scan per
    dirty-scheduler call, nifs.rs:297-309). Honours --steps / --warmup; when the whole run would exceed
    ~2.5 minutes the per-step corpus sample shrinks (stated in `sample`) and the rate is scaled linearly."""
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
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    t0 = time.perf_counter()
    oracle.flat_scan_timed("cosine", rows, queries[:threads], args.k, threads)   # page in + first estimate
    est = time.perf_counter() - t0
    budget = 150.0
    sample_rows = args.rows
    if (steps + warmup) * est > budget:
        sample_rows = max(10_000, int(args.rows * budget / ((steps + warmup) * est)))
    sample = rows[:sample_rows]
    for _ in range(warmup):
        oracle.flat_scan_timed("cosine", sample, queries[:threads], args.k, threads)
    t = 0.0
    for s in range(steps):
        q = np.roll(queries, s, axis=0)[:threads]
        secs, _ = oracle.flat_scan_timed("cosine", sample, q, args.k, threads)
        t += secs
    qps = steps * threads / t * (sample_rows / args.rows)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.rows, args.dim, args.k),
                   "note": "CPU restatement of flat.rs:96-124 (oracle port); one step = one query per host thread"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x {threads} concurrent single-query scans of "
                                   + (f"the full {args.rows}x{args.dim} corpus" if sample_rows == args.rows else
                                      f"the first {sample_rows} of {args.rows} rows x {args.dim} (rate scaled linearly "
                                      f"to the full corpus)")},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- CUDA arm
class Run:
    """Process-group plumbing shared by the headline and the config blocks."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def dev_timed(self, fn, steps: int, warm: int = 2) -> float:
        """ms per call: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.allmax(e0.elapsed_time(e1) / steps)

    def wall_timed(self, fn, steps: int, warm: int = 2) -> float:
        """ms per call by wall clock (for calls that synchronise internally), max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.allmax((time.perf_counter() - t0) / steps * 1e3)

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def ingest_rows_device(run: Run, index, rows: int, dim: int, base: int, seed: int, keep_first: int = 0):
    """Generates `rows` normalised rows on the device chunk by chunk and hands them to the index through the
    device bulk-ingest entry (vb_flat_insert_many_device); ids are zero-padded global row numbers. Returns
    (seconds, host copy of the first `keep_first` rows for the parity checks)."""
    from vettore_b200 import nifs

    torch = run.torch
    assert nifs.flat_reserve(index, rows) == ("ok", ())
    kept = None
    chunk = INGEST_CHUNK
    t0 = time.perf_counter()
    for s in range(0, rows, chunk):
        m = min(chunk, rows - s)
        blk = make_rows_torch(chunk, dim, seed + 7 * (s // chunk), run.dev)[:m]   # always a whole chunk: RowRegen re-creates it
        res = nifs.flat_insert_device(index, nifs.decimal_ids(base + s, m), blk.data_ptr(), dim)
        assert res == ("ok", ()), res
        if s == 0 and keep_first:
            kept = blk[:keep_first].cpu().numpy()
        del blk
    torch.cuda.synchronize()
    return time.perf_counter() - t0, kept


class RowRegen:
    """Re-creates any row of a device-generated corpus on the host for the oracle: the corpora of configs 3-5
    are generated chunk by chunk from per-chunk seeds (ingest_rows_device), so the chunk holding a given row
    can be generated again on the device and the row fetched. One chunk is cached."""

    def __init__(self, run: Run, dim: int, seed_of_chunk, chunk_rows: int):
        self.run, self.dim, self.seed_of_chunk, self.chunk_rows = run, dim, seed_of_chunk, chunk_rows
        self.cached, self.block = None, None

    def rows(self, shard: int, row0: int, count: int = 1):
        """`count` consecutive rows starting at local row `row0` of shard `shard` (all inside one chunk)."""
        c = row0 // self.chunk_rows
        if self.cached != (shard, c):
            self.block = None
            self.run.free()
            self.block = make_rows_torch(self.chunk_rows, self.dim, self.seed_of_chunk(shard, c), self.run.dev)
            self.cached = (shard, c)
        o = row0 - c * self.chunk_rows
        return self.block[o:o + count].cpu().numpy()

    def close(self):
        self.block, self.cached = None, None
        self.run.free()


def headline(run: Run, args, pk):
    import numpy as np

    import oracle  # parity checks only
    from vettore_b200 import nifs
    from vettore_b200._lib import lib
    from vettore_b200.sharded import ShardedFlat, set_global_ranks

    torch, world, rank, dev = run.torch, run.world, run.rank, run.dev
    n, d, k = args.rows, args.dim, args.k
    # ---- corpus shard: generated on the device, brought to the host once, then ingested through the
    # reference-facing boundary (flat_insert_many -> bulk H2D).
    x = make_rows_torch(n, d, SEED + 7919 * rank, dev)
    host_rows = x.cpu().numpy()
    del x
    run.free()
    base = rank * n
    ids = nifs.decimal_ids(base, n)
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

    # ---- parity against the oracle BEFORE timing, on the benchmarked configuration itself: rank 0 runs
    # every rotating query through the oracle over its whole shard (all host threads) and compares with
    # the CUDA path's hits through the reference-facing call. (The timed CPU baseline reuses this scan.)
    threads = os.cpu_count() or 1
    parity_checked, cpu_secs, cpu_nq = 0, None, 0
    local_hits = []
    for qi in range(args.queries):
        st, hits = nifs.flat_search(index, q_host[qi].numpy(), k)
        assert st == "ok", hits
        local_hits.append([(int(h[0]) - base, h[1]) for h in hits])
    if rank == 0:
        cpu_nq = args.queries if world == 1 else min(args.queries, 16)
        cpu_secs, ref = oracle.flat_scan_timed("cosine", host_rows, q_host[:cpu_nq].numpy(), k, threads)
        for qi in range(cpu_nq):
            assert_hits_match(local_hits[qi], ref[qi], f"headline query {qi}")
        parity_checked = cpu_nq
    if world > 1:
        # the merged global answer must be the merge of the (oracle-checked) per-shard answers
        nchk = 8
        mine = [[(run.rank, r, v) for r, v in local_hits[qi]] for qi in range(nchk)]
        gathered = [None] * world
        run.dist.all_gather_object(gathered, mine)
        for qi in range(nchk):
            got = sharded.search(q_host[qi:qi + 1])[0]
            pool = sorted((e for g in gathered for e in g[qi]), key=lambda e: (np.float32(1.0) - np.float32(e[2]), e[0], e[1]))
            exp = pool[:k]
            assert [(h.shard, h.row) for h in got] == [(e[0], e[1]) for e in exp], ("global merge", qi)

    # ---- value: device-timed, query already resident in HBM
    def step_device(i):
        return sharded.search_device(queries[i % args.queries: i % args.queries + 1])

    for i in range(args.warmup):
        step_device(i)
    run.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(run.local_rank) as clocks:
        ev0.record()
        for i in range(args.steps):
            step_device(i)
        ev1.record()
        run.barrier()
    ms_total = run.allmax(ev0.elapsed_time(ev1))

    # ---- scan-kernel-only timing (roofline): the local scan without the exchange
    lay = sharded.layout

    def scan_only(i):
        q = queries[i % args.queries: i % args.queries + 1]
        rc = lib().vb_flat_search_device(index.handle, C.c_void_p(q.data_ptr()), 1, d, k,
                                         C.c_void_p(sharded.local.data_ptr() + lay["keys"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["values"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["rows"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["counts"]), run.stream)
        assert rc == 0

    run.barrier()
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
    run.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    run.barrier()
    e2e_s = run.allmax(time.perf_counter() - t0)

    # ---- configs[1] also names a batch of 1024 queries: K2 (tcgen05 3xTF32 GEMM + fused top-k), device-timed
    # with the queries resident in HBM, parity-checked against the oracle on 32 of its queries.
    batch = None
    if world == 1 and "c2" in args.cfgs:
        nqb = 1024
        bq = make_rows_torch(nqb, d, SEED + 2, dev)
        bkeys = torch.zeros(nqb * k, dtype=torch.int64, device=dev)
        bvals = torch.zeros(nqb * k, dtype=torch.float32, device=dev)
        brows = torch.zeros(nqb * k, dtype=torch.int32, device=dev)
        bcnts = torch.zeros(nqb, dtype=torch.int32, device=dev)

        def batch_step():
            rc = lib().vb_flat_search_device(index.handle, C.c_void_p(bq.data_ptr()), nqb, d, k,
                                             C.c_void_p(bkeys.data_ptr()), C.c_void_p(bvals.data_ptr()),
                                             C.c_void_p(brows.data_ptr()), C.c_void_p(bcnts.data_ptr()), run.stream)
            assert rc == 0

        bms = run.dev_timed(batch_step, 5, 2)
        nchk = 32
        bq_host = bq[:nchk].cpu().numpy()
        _, bref = oracle.flat_scan_timed("cosine", host_rows, bq_host, k, threads)
        rows_h, vals_h = brows.cpu().numpy().reshape(nqb, k), bvals.cpu().numpy().reshape(nqb, k)
        for qi in range(nchk):
            assert_hits_match([(int(rows_h[qi, i]), float(vals_h[qi, i])) for i in range(k)], bref[qi], f"batch query {qi}")
        tf32_peak = pk["tf32"]
        terms = int(lib().vb_debug_gemm_terms())   # TF32 passes of the last batch (3 = it fell back to the 3xTF32 tier)
        alg = 2.0 * nqb * n * d / (bms * 1e-3) / 1e12
        issued = terms * alg
        batch = {"workload": f"flat cosine exact scan {n}x{d} fp32, batch of {nqb} queries, k={k}",
                 "queries": nqb, "ms_per_batch": bms, "queries_per_sec": nqb / (bms * 1e-3),
                 "tf32_passes": terms, "tf32_tflops_issued": issued, "algorithmic_tflops": alg,
                 "roofline": {"bound": "tensor", "achieved": issued, "peak": tf32_peak, "unit": "TFLOP/s",
                              "frac": issued / tf32_peak, "peak_source": pk["tc_src"],
                              "note": "the tensor-core pass is a candidate filter in front of the exact fp32 re-scoring: ONE "
                                      "TF32 pass (issued = algorithmic flops), 3xTF32 only as the fallback tier for queries whose "
                                      "candidate set could not be proven complete; the sample pre-pass, list merges and "
                                      "re-scoring launches are inside ms_per_batch"},
                 "parity_checked": nchk,
                 "kernel": "vb::flat_gemm1_topk_kernel (tcgen05 kind::tf32, SS mode, fused top-k filter) + exact re-scoring"
                           if terms == 1 else "vb::flat_gemm_topk_kernel (tcgen05 3xTF32) + exact re-scoring"}

    line = None
    if rank == 0:
        alg_bytes = n * d * 4
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC,
            "value": world * args.steps / (ms_total * 1e-3),
            "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "parity_checked": parity_checked,
            "config": {
                "workload": workload_string(n, d, k),
                "value_definition": "1M-row shard scans per second over all GPUs = (queries/s over the "
                                    f"{n * world}-row sharded corpus) x n_gpus; at n_gpus=1 plain queries/s",
                "corpus_rows_total": n * world, "parallelism": f"row-shard x{world}",
                "exchange": sharded.exchange_name,
                "l2_policy": f"corpus shard {alg_bytes / 1e9:.2f} GB >> 126 MB L2, {args.queries} rotating queries",
                "queries_per_sec": args.steps / (ms_total * 1e-3),
                "ingest_seconds_per_shard": round(ingest_s, 3),
                "parity": f"{parity_checked} of the rotating queries: CUDA hits == oracle hits over the whole shard "
                          "(ids equal, values within 1e-5)" + ("; merged global top-k == merge of the per-shard lists"
                                                              if world > 1 else ""),
            },
            "clocks": clocks.summary(),
            "e2e": {"value": world * args.steps / e2e_s, "unit": "queries/s",
                    "h2d_bytes_per_step": d * 4, "d2h_bytes_per_step": k * 8 + 8 if world == 1 else int(sharded.out.numel()),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": args.steps * sharded.launches_per_search,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": achieved / pk["hbm"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full
                         # capture of this shape (profiles/r1_flat_stream_ncu_summary.txt: 3.073186 GB + 6.35 MB)
                         "traffic": 3.078051e9 if (n, d, k) == (1_000_000, 768, 10) else None,
                         "traffic_source": "profiles/r2_flat_stream_ncu_summary.txt (ncu --set full of this kernel inside this command, "
                                           "round 2: dram__bytes_read 3.073062 GB + dram__bytes_write 4.99 MB per launch)",
                         "peak_source": pk["hbm_src"],
                         "kernel": "vb::flat_stream_kernel<cosine, NV=6, RPW=1, W=16> (TMA-staged ring; the timed "
                                   "launch pair also holds the ~3 us unpack kernel)",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms},
        }
        if batch is not None:
            line["config"]["c2_batch_1024"] = batch
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": cpu_nq / cpu_secs, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": f"{cpu_nq} single-query scans of the same {n}x{d} corpus, {threads} host "
                                              f"threads, one sequential scan per query (flat.rs:96-124 restated); the "
                                              f"same scan is the parity check"}
            # the reference runs ONE sequential scan per flat_search call (flat.rs:104-118): the 1-thread figure
            s1, _ = oracle.flat_scan_timed("cosine", host_rows, q_host[:2].numpy(), k, 1)
            line["cpu_baseline"]["one_thread"] = {"value": 2 / s1, "unit": "queries/s", "cores": 1,
                                                  "sample": f"2 sequential scans of the same {n}x{d} corpus on one thread"}
    del sharded, index
    run.free()
    return line


# ------------------------------------------------------------------------------- BASELINE configs 3-5
def rows_per_gpu(total: int, world: int, cap: int) -> int:
    return min(total // world, cap)


def block_c3(run: Run, args, pk):
    """configs[2]: flat inner product, 100M x 768 row-sharded, 1024-query batch, k=100."""
    import numpy as np

    import oracle
    from vettore_b200 import nifs
    from vettore_b200._lib import lib
    from vettore_b200.sharded import ShardedFlat, set_global_ranks

    torch, world, rank, dev = run.torch, run.world, run.rank, run.dev
    total, d, nq, k = int(100_000_000 * args.corpus_scale), 768, 1024, 100
    free_b, _ = torch.cuda.mem_get_info()
    cap_rows = int((free_b - 12e9) // (d * 4 + 4))                 # matrix + ranks, 12 GB left for workspaces
    want = total // world if world > 1 else total // 8             # one GPU: the G=8 shard (the full corpus is 307 GB)
    n = min(want, cap_rows)
    if world > 1 and n < want:
        n = min(n, int(25_000_000 * args.corpus_scale))            # does not fit: a 25M-row shard per GPU instead
    base = rank * n
    index = nifs.flat_new_inner_product()
    ingest_s, kept = ingest_rows_device(run, index, n, d, base, SEED + 1000 * rank, keep_first=100_000 if rank == 0 else 0)
    if world > 1:
        set_global_ranks(index, base, n)
    queries = make_rows_torch(nq, d, SEED + 3, dev)
    sharded = ShardedFlat(index, k=k, nq=nq, device=dev)
    lay = sharded.layout

    def local_only():
        rc = lib().vb_flat_search_device(index.handle, C.c_void_p(queries.data_ptr()), nq, d, k,
                                         C.c_void_p(sharded.local.data_ptr() + lay["keys"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["values"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["rows"]),
                                         C.c_void_p(sharded.local.data_ptr() + lay["counts"]), run.stream)
        assert rc == 0

    with ClockSampler(run.local_rank) as c3_clocks:
        local_ms = run.dev_timed(local_only, 5, 1)
    step_ms = run.dev_timed(lambda: sharded.search_device(queries), 3, 1) if world > 1 else local_ms
    out_host = sharded.search_device(queries).cpu().numpy()
    hits = sharded.decode(out_host, nq)
    checked = 0
    if rank == 0:
        nchk = 4
        _, ref = oracle.flat_scan_timed("inner_product", kept, queries[:nchk].cpu().numpy(), k, os.cpu_count() or 1)
        for qi in range(nchk):
            got = [(h.row if h.shard == 0 else 1 << 40, h.value) for h in hits[qi]]
            checked += subsample_check(got, ref[qi], kept.shape[0], True, f"c3 query {qi}")
        # every returned value must be the oracle's inner product of the query with that very row: the best 8
        # hits of each checked query, rows re-created from their chunk seeds (any shard)
        regen = RowRegen(run, d, lambda shard, c: SEED + 1000 * shard + 7 * c, INGEST_CHUNK)
        qh = queries[:nchk].cpu().numpy()
        for qi in range(nchk):
            for h in hits[qi][:8]:
                st, v = oracle.compute("inner_product", qh[qi], regen.rows(h.shard, h.row)[0])
                assert st == "ok" and close(h.value, v), ("c3 value", qi, h.shard, h.row, h.value, v)
                checked += 1
        regen.close()
    terms = int(lib().vb_debug_gemm_terms())   # 1: single TF32 filter pass; 3: the batch fell back to the 3xTF32 tier
    issued = terms * 2.0 * nq * n * d / (local_ms * 1e-3) / 1e12
    tf32_peak = pk["tf32"]
    out = {"workload": f"flat inner-product scan {total}x{d} fp32 row-sharded x{world}, batch of {nq} queries, k={k}",
           "rows_per_gpu": n, "corpus_rows_resident": n * world, "full_corpus": n * world == total,
           "fits_note": None if n * world == total else
           (f"one GPU holds the G=8 shard ({n} rows): the full corpus is {total * d * 4 / 1e9:.0f} GB" if world == 1 else
            f"{total // world} rows x {d} fp32 = {total // world * d * 4 / 1e9:.0f} GB per GPU does not fit next to the "
            f"workspaces: {n} rows per GPU resident"),
           "local_scan_ms": local_ms, "step_ms": step_ms, "queries_per_sec": nq / (step_ms * 1e-3),
           "tf32_passes": terms, "per_gpu_tf32_tflops_issued": issued, "per_gpu_algorithmic_tflops": issued / terms,
           "kernel": ("vb::flat_gemm1_topk_kernel (one TF32 pass as a candidate filter, k' = 192 kept per query)" if terms == 1
                      else "vb::flat_gemm_topk_kernel (3xTF32)") + " + exact fp32 re-scoring",
           "roofline": {"bound": "tensor", "achieved": issued, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": issued / tf32_peak, "peak_source": pk["tc_src"]},
           "speedup_vs_one_gpu_same_corpus": world * local_ms / step_ms,
           "speedup_definition": "one GPU works through the same resident corpus shard by shard: n_gpus x local_scan_ms / step_ms",
           "exchange_bytes_per_rank": lay["bytes"], "exchange": sharded.exchange_name,
           "clocks_during_local_scan": c3_clocks.summary(),
           "ingest_seconds_per_shard": round(ingest_s, 2), "ingest_rows_per_sec": n / ingest_s,
           "parity": f"4 queries: order + completeness vs the oracle's top-{k} over the first {0 if kept is None else kept.shape[0]} "
                     f"rows of shard 0; the best 8 hits of each re-scored by the oracle on rows re-created from their "
                     f"chunk seeds ({checked} values verified within 1e-5)"}
    del sharded, index, queries
    run.free()
    return out


def block_c4(run: Run, args, pk):
    """configs[3]: quantized_search, 100M x 1024-bit sign codes, 1000 candidates -> exact cosine rerank to 10."""
    import numpy as np

    import oracle
    from vettore_b200 import nifs
    from vettore_b200._lib import lib
    from vettore_b200.sharded import ShardedQuantized, set_global_ranks

    torch, world, rank, dev = run.torch, run.world, run.rank, run.dev
    total, d, cand, k = int(100_000_000 * args.corpus_scale), 1024, 1000, 10
    free_b, _ = torch.cuda.mem_get_info()
    cap_rows = int((free_b - 12e9) // (d * 4 + d // 8 + 4))         # fp32 rows (rerank) + codes + ranks
    want = total // world if world > 1 else total // 8
    n = min(want, cap_rows)
    if world > 1 and n < want:
        n = min(n, int(25_000_000 * args.corpus_scale))
    base = rank * n
    index = nifs.flat_new_cosine()
    keep = 200_000 if rank == 0 else 0
    ingest_s, kept = ingest_rows_device(run, index, n, d, base, SEED + 2000 * rank, keep_first=keep)
    if world > 1:
        set_global_ranks(index, base, n)
    queries = make_rows_torch(8, d, SEED + 4, dev)
    sq = ShardedQuantized(index, cand, k, nifs.METRIC_CODE["cosine"], device=dev)
    lc = sq.lay_c
    q0 = queries[0:1].contiguous()

    def hamming_local():
        rc = lib().vb_flat_hamming_device(index.handle, C.c_void_p(q0.data_ptr()), 1, d, cand,
                                          C.c_void_p(sq.local_c.data_ptr() + lc["keys"]),
                                          C.c_void_p(sq.local_c.data_ptr() + lc["values"]),
                                          C.c_void_p(sq.local_c.data_ptr() + lc["rows"]),
                                          C.c_void_p(sq.local_c.data_ptr() + lc["counts"]), run.stream)
        assert rc == 0

    hamming_local()                                     # builds the code mirror (K6) outside the timed region
    torch.cuda.synchronize()
    ham_local_ms = run.dev_timed(hamming_local, 20, 3)
    ham_step_ms = run.dev_timed(lambda: sq.candidates_device(q0), 20, 3) if world > 1 else ham_local_ms
    pipe_ms = run.dev_timed(lambda: sq.search_device(q0), 20, 3)
    # parity: Hamming candidates bit-exact on the subsample property; rerank values vs the oracle
    checked = 0
    cand_rec, offs_c = sq.candidates_device(q0)
    cand_host = cand_rec.cpu().numpy()
    from vettore_b200.sharded import _decode_merged
    cands = _decode_merged(cand_host, offs_c, cand)
    final = sq.search(queries[0:1].cpu().pin_memory())
    if rank == 0:
        qh = queries[0].cpu().numpy()
        qbits = np.array(oracle.compress_sign_bits(qh), dtype=np.uint64)
        codes = np.array([oracle.compress_sign_bits(r) for r in kept[:50_000]], dtype=np.uint64)
        _, ref = oracle.binary_scan_timed(codes, d, qbits[None, :], cand, os.cpu_count() or 1)
        got = [(h.row if h.shard == 0 else 1 << 40, h.value) for h in cands]
        # integer path: values must be EQUAL for the rows inside the sample
        refd = dict(ref[0])
        for r, v in got:
            if r < codes.shape[0] and r in refd:
                assert v == refd[r], ("c4 hamming distance", r, v, refd[r])
                checked += 1
        worst = got[-1][1]
        got_rows = {r for r, _ in got}
        for r, v in ref[0]:
            if v < worst:
                assert r in got_rows, ("c4: oracle candidate missing", r, v, worst)
        # the final hits must be candidates, sorted, and carry the oracle's f64 cosine for rows in the sample
        cset = {(h.shard, h.row) for h in cands}
        assert all((h.shard, h.row) in cset for h in final), "c4: final hit outside the candidate set"
        regen = RowRegen(run, d, lambda shard, c: SEED + 2000 * shard + 7 * c, INGEST_CHUNK)
        for h in final:             # every final hit: the oracle's f64 cosine of the query with that very row
            st, v = oracle.cosine(qh, regen.rows(h.shard, h.row)[0])
            assert st == "ok" and close(h.value, v), ("c4 rerank value", h.shard, h.row, h.value, v)
            checked += 1
        for h in cands[:8] + cands[-8:]:   # best and worst candidates: exact Hamming distance of the sign codes
            bits = np.array(oracle.compress_sign_bits(regen.rows(h.shard, h.row)[0]), dtype=np.uint64)
            st, dist = oracle.packed_hamming(qbits, bits, d)
            assert st == "ok" and dist == h.value, ("c4 candidate distance", h.shard, h.row, h.value, dist)
            checked += 1
        regen.close()
    code_bytes = n * (d // 64) * 8
    gbs = code_bytes / (ham_local_ms * 1e-3) / 1e9
    out = {"workload": f"quantized_search: {total}x{d}-bit sign codes row-sharded x{world}, {cand} candidates -> exact cosine rerank to k={k}",
           "rows_per_gpu": n, "corpus_rows_resident": n * world, "full_corpus": n * world == total,
           "fits_note": None if n * world == total else
           (f"one GPU holds the G=8 shard ({n} rows): the fp32 rows the rerank needs are {total * d * 4 / 1e9:.0f} GB in all"
            if world == 1 else f"{n} rows per GPU resident (fp32 rows for the rerank + codes)"),
           "hamming_pass": {"local_scan_ms": ham_local_ms, "step_ms": ham_step_ms,
                            "queries_per_sec": 1e3 / ham_step_ms, "per_gpu_gbs": gbs,
                            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                                         "frac": gbs / pk["hbm"], "peak_source": pk["hbm_src"],
                                         "algorithmic_bytes_per_launch": code_bytes},
                            "speedup_vs_one_gpu_same_corpus": world * ham_local_ms / ham_step_ms},
           "pipeline": {"step_ms": pipe_ms, "queries_per_sec": 1e3 / pipe_ms,
                        "stages": "K6 sign-pack(query) + K3 Hamming scan + exchange/select + K4 exact rerank of the owned "
                                  "candidates (one stream sync: the owned count sizes the launch) + exchange/select"},
           "speedup_definition": "Hamming pass: n_gpus x local_scan_ms / step_ms (one GPU works through the same codes shard by shard)",
           "exchange": sq.exchange_name if world > 1 else "none",
           "ingest_seconds_per_shard": round(ingest_s, 2),
           "parity": f"candidates: Hamming distances EQUAL to the oracle's on a 50k-row sample of shard 0, no better oracle "
                     f"candidate missing, best/worst 8 re-derived from re-created rows; all {k} final hits are candidates and "
                     f"carry the oracle's f64 cosine ({checked} values verified)"}
    del sq, index, queries
    run.free()
    return out


def block_c5(run: Run, args, pk):
    """configs[4]: ColBERT MaxSim, 1M docs x 128 tokens x 128 dims, 32-token query, k=10, document-sharded."""
    import numpy as np

    import oracle
    from vettore_b200 import nifs
    from vettore_b200.sharded import ShardedMv, set_global_mv_ranks

    torch, world, rank, dev = run.torch, run.world, run.rank, run.dev
    total, td, d, tq, k = int(1_000_000 * args.corpus_scale), 128, 128, 32, 10
    n = total // world                                              # 65.5 GB in all: fits one GPU
    base = rank * n
    index = nifs.mv_new("inner_product")
    assert nifs.mv_reserve(index, n, n * td, d) == ("ok", ())
    chunk = 20_000
    kept = None
    t0 = time.perf_counter()
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        x = make_rows_torch(chunk * td, d, SEED + 3000 * rank + 7 * (s // chunk), dev)[:m * td]   # whole chunks: RowRegen
        assert nifs.mv_insert_device(index, nifs.decimal_ids(base + s, m), x.data_ptr(), td, d) == ("ok", ())
        if s == 0 and rank == 0:
            kept = x[: 2000 * td].cpu().numpy().reshape(-1, td, d)
        del x
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t0
    if world > 1:
        set_global_mv_ranks(index, base, n)
    q = make_rows_torch(tq, d, SEED + 5, dev).cpu().numpy()
    smv = ShardedMv(index, k, device=dev)
    with ClockSampler(run.local_rank) as c5_clocks:            # the 65 GB scan runs under combined tensor + HBM load
        local_ms = run.wall_timed(lambda: nifs.mv_search(index, q, k), 20 if world == 1 else 40, 2)
    step_ms = run.wall_timed(lambda: smv.search(q), 10, 2) if world > 1 else local_ms
    hits = smv.search(q)
    checked = 0
    if rank == 0:
        _, ref = oracle.maxsim_scan_timed("inner_product", kept, q[None, :, :], k, os.cpu_count() or 1)
        got = [(h.row if h.shard == 0 else 1 << 40, h.value) for h in hits]
        checked = subsample_check(got, ref[0], kept.shape[0], True, "c5")
        # every returned score must be the oracle's MaxSim of the query with that very document (any shard),
        # its tokens re-created from the chunk seed
        regen = RowRegen(run, d, lambda shard, c: SEED + 3000 * shard + 7 * c, chunk * td)
        for h in hits:
            st, v = oracle.multi_vector_score(q, regen.rows(h.shard, h.row * td, td), nifs.METRIC_CODE["inner_product"])
            assert st == "ok" and close(h.value, v), ("c5 score", h.shard, h.row, h.value, v)
            checked += 1
        regen.close()
    alg = n * td * d * 4
    gbs = alg / (local_ms * 1e-3) / 1e9
    out = {"workload": f"multi_vector_search MaxSim: {total} docs x {td} tokens x {d} dims fp32 document-sharded x{world}, "
                       f"{tq}-token query, k={k}",
           "docs_per_gpu": n, "full_corpus": True,
           "local_scan_ms": local_ms, "step_ms": step_ms, "queries_per_sec": 1e3 / step_ms, "per_gpu_gbs": gbs,
           "per_gpu_algorithmic_tflops": 2.0 * tq * n * td * d / (local_ms * 1e-3) / 1e12,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                        "peak_source": pk["hbm_src"], "algorithmic_bytes_per_launch": alg},
           "timing": "wall clock through the by-value query call (vb_mv_search synchronises): H2D of the query, K5, D2H of the hits",
           "strong_scaling": "the 1M-document corpus is fixed; docs_per_gpu = 1M / n_gpus (compare queries_per_sec across the N lines)",
           "speedup_vs_one_gpu_same_corpus": world * local_ms / step_ms,
           "exchange": smv.exchange_name, "phase_ms": smv.phase_ms(),
           "clocks_during_local_scan": c5_clocks.summary(),
           "ingest_seconds_per_shard": round(ingest_s, 2),
           "parity": f"order + completeness vs the oracle's top-{k} over the first {kept.shape[0] if kept is not None else 0} "
                     f"documents of shard 0; all {k} returned scores re-derived by the oracle from documents re-created from "
                     f"their chunk seeds ({checked} values verified within 1e-5)"}
    del smv, index
    run.free()
    if world == 1:
        out["ragged"] = block_c5_ragged(run, args, pk)
    return out


def block_c5_ragged(run: Run, args, pk):
    """The C5 shape with RAGGED documents (40-180 tokens, the length profile of a late-interaction corpus): the ragged
    tensor-core kernel (csrc/maxsim_tcr.cu), one GPU, device ingest, parity against the oracle on the first documents."""
    import numpy as np

    import oracle
    from vettore_b200 import _lib, nifs

    torch, dev = run.torch, run.dev
    ndocs, d, tq, k = int(200_000 * args.corpus_scale), 128, 32, 10
    lens = np.random.default_rng(SEED).integers(40, 181, ndocs)
    total = int(lens.sum())
    index = nifs.mv_new("inner_product")
    assert nifs.mv_reserve(index, ndocs, total, d) == ("ok", ())
    chunk, kept = 20_000, None
    for s in range(0, ndocs, chunk):
        m = min(chunk, ndocs - s)
        doc_tok = np.concatenate([[0], np.cumsum(lens[s:s + m])]).astype(np.uint64)
        x = make_rows_torch(int(doc_tok[-1]), d, SEED + 11 * (s // chunk), dev)
        assert nifs.mv_insert_ragged_device(index, nifs.decimal_ids(s, m), x.data_ptr(), doc_tok, d) == ("ok", ())
        if s == 0:
            c = min(2000, m)
            host = x[: int(doc_tok[c])].cpu().numpy()
            kept = [(f"{i:09d}", host[int(doc_tok[i]):int(doc_tok[i + 1])]) for i in range(c)]
        del x
    torch.cuda.synchronize()
    q = make_rows_torch(tq, d, SEED + 5, dev).cpu().numpy()
    ms = run.wall_timed(lambda: nifs.mv_search(index, q, k), 10, 2)
    st, hits = nifs.mv_search(index, q, k)
    assert st == "ok", hits
    path = int(_lib.lib().vb_debug_maxsim_path())
    ref = dict(oracle.multi_vector_top_k(kept, q, nifs.METRIC_CODE["inner_product"], len(kept))[1])
    checked = 0
    for hid, v in hits:                      # every hit that lies in the kept prefix carries the oracle's score ...
        if hid in ref:
            assert close(v, ref[hid]), ("c5 ragged score", hid, v, ref[hid])
            checked += 1
    for hid, v in hits:                      # ... and EVERY returned score is the oracle's MaxSim of that very document,
        j = int(hid)                         # its tokens re-created from the chunk seed
        s0 = (j // chunk) * chunk
        off = np.concatenate([[0], np.cumsum(lens[s0:s0 + chunk])])
        x = make_rows_torch(int(off[-1]), d, SEED + 11 * (s0 // chunk), dev)
        doc = x[int(off[j - s0]):int(off[j - s0 + 1])].cpu().numpy()
        del x
        st, sc = oracle.multi_vector_score(q, doc, nifs.METRIC_CODE["inner_product"])
        assert st == "ok" and close(v, sc), ("c5 ragged re-derived score", hid, v, sc)
        checked += 1
    worst = hits[-1][1]                      # ... and no kept document beats the k-th hit without being a hit
    got_ids = {h[0] for h in hits}
    assert all(v <= worst or i in got_ids or close(v, worst) for i, v in ref.items()), "c5 ragged completeness"
    alg = total * d * 4
    gbs = alg / (ms * 1e-3) / 1e9
    del index
    run.free()
    return {"workload": f"multi_vector_search MaxSim over RAGGED documents: {ndocs} docs of 40-180 tokens ({total} tokens) x {d} dims, "
                        f"{tq}-token query, k={k}, one GPU",
            "kernel": {2: "vb::maxsim_tcr_kernel (tcgen05 3xTF32, segmented max over document boundaries)", 1: "uniform tensor-core kernel",
                       0: "general CUDA-core kernel"}.get(path, str(path)),
            "step_ms": ms, "queries_per_sec": 1e3 / ms, "per_gpu_gbs": gbs,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                         "peak_source": pk["hbm_src"], "algorithmic_bytes_per_launch": alg},
            "parity": f"all {k} returned scores re-derived by the oracle from documents re-created from their chunk seeds "
                      f"({checked} values verified within 1e-5); completeness against the oracle over the first {len(kept)} documents"}


def cpu_samples(args):
    """Bounded CPU legs for configs 3-5 (rank 0, N=1): the oracle on a sample of each shape, all host threads,
    rates scaled linearly to the stated corpus (shapes too large for host RAM are extrapolated)."""
    import numpy as np
    import torch

    import oracle

    threads = os.cpu_count() or 1
    cpu = torch.device("cpu")
    out = {"cores": threads, "kind": "port"}
    rows = make_rows_torch(200_000, 768, SEED, cpu).numpy()
    q = make_rows_torch(threads, 768, SEED + 1, cpu).numpy()
    s, _ = oracle.flat_scan_timed("inner_product", rows, q, 100, threads)
    out["c3"] = {"queries_per_sec_100M_rows": threads / s * (200_000 / 100_000_000),
                 "sample": f"{threads} queries x 200k x 768 rows, k=100, scaled linearly to 100M rows"}
    rng = np.random.default_rng(SEED)
    codes = rng.integers(0, 2 ** 63, size=(2_000_000, 16), dtype=np.uint64)
    qb = rng.integers(0, 2 ** 63, size=(threads, 16), dtype=np.uint64)
    s, _ = oracle.binary_scan_timed(codes, 1024, qb, 1000, threads)
    out["c4_hamming_pass"] = {"queries_per_sec_100M_rows": threads / s * (2_000_000 / 100_000_000),
                              "sample": f"{threads} queries x 2M x 1024-bit codes, 1000 candidates, scaled linearly to 100M rows"}
    toks = make_rows_torch(500 * 128, 128, SEED + 2, cpu).numpy().reshape(500, 128, 128)
    qt = make_rows_torch(threads * 32, 128, SEED + 3, cpu).numpy().reshape(threads, 32, 128)
    s, _ = oracle.maxsim_scan_timed("inner_product", toks, qt, 10, threads)
    out["c5"] = {"queries_per_sec_1M_docs": threads / s * (500 / 1_000_000),
                 "sample": f"{threads} queries x 500 docs x 128 x 128, scaled linearly to 1M docs"}
    return out


def run_b200(args):
    run = Run()
    pk = peaks()
    tf32 = measured_tf32_peak()
    if tf32:   # the tensor-bound rooflines are stated against the pipe's own measured TF32 rate
        pk["tf32"], pk["tc_src"] = tf32, "measured on this device: bare tcgen05.mma kind::tf32 loop (vb_debug_tf32_peak)"
    else:
        pk["tf32"] = pk["bf16"] / 2
    args.cfgs = [c for c in args.configs.split(",") if c]
    line = headline(run, args, pk)
    blocks = {}
    for name, fn in (("c3", block_c3), ("c4", block_c4), ("c5", block_c5)):
        if name not in args.cfgs:
            continue
        try:
            t0 = time.perf_counter()
            blk = fn(run, args, pk)
            blk["block_seconds"] = round(time.perf_counter() - t0, 1)
            blocks[name] = blk
        except Exception as e:  # a config block must never take the headline line down with it
            import traceback
            blocks[name] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}
            run.free()
    if run.rank == 0:
        line["config"].update(blocks)
        if run.world == 1 and not args.no_cpu_baseline and blocks:
            try:
                line["config"]["cpu_baseline_c3_c5"] = cpu_samples(args)
            except Exception as e:
                line["config"]["cpu_baseline_c3_c5"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if run.world > 1:
        run.barrier()
        run.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
