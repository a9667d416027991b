#!/usr/bin/env python
"""Secondary benchmark: ColBERT MaxSim (BASELINE.json configs[4] shape per GPU: docs x 128 tokens
x 128 dims fp32, 32-token query, k=10) on the HBM-resident multi-vector index.
Prints one JSON line with queries/s, achieved GB/s (tokens*D*4 bytes per query) and the
fraction of the measured HBM peak. --general forces the CUDA-core kernel for comparison."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=200_000)
    ap.add_argument("--tokens", type=int, default=128)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--tq", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--metric", default="inner_product")
    ap.add_argument("--general", action="store_true")
    ap.add_argument("--ragged", default=None, metavar="LO,HI",
                    help="document lengths uniform in [LO, HI] (ragged tensor-core kernel); --docs counts documents")
    ap.add_argument("--check", type=int, default=2000, help="docs re-scored by the oracle for a parity spot check")
    args = ap.parse_args()
    if args.general:
        os.environ["VB_MAXSIM_NO_TC"] = "1"
    import numpy as np
    import torch

    from vettore_b200 import _lib, nifs

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev)
    g.manual_seed(20_260_721)
    idx = nifs.mv_new(args.metric)
    lens = None
    if args.ragged:
        lo, hi = (int(v) for v in args.ragged.split(","))
        lens = np.random.default_rng(7).integers(lo, hi + 1, args.docs)
    total_tokens = int(lens.sum()) if lens is not None else args.docs * args.tokens
    assert nifs.mv_reserve(idx, args.docs, total_tokens, args.dim) == ("ok", ())
    chunk = 20_000
    first = None
    t0 = time.perf_counter()
    for s in range(0, args.docs, chunk):
        n = min(chunk, args.docs - s)
        rows = int(lens[s:s + n].sum()) if lens is not None else n * args.tokens
        x = torch.randn(rows, args.dim, generator=g, device=dev)
        x = (x.double() / x.double().norm(dim=1, keepdim=True)).float().contiguous()
        ids = [f"{s + i:09d}" for i in range(n)]
        if lens is not None:
            doc_tok = np.concatenate([[0], np.cumsum(lens[s:s + n])]).astype(np.uint64)
            assert nifs.mv_insert_ragged_device(idx, ids, x.data_ptr(), doc_tok, args.dim) == ("ok", ())
            if first is None:
                c = min(args.check, n)
                host = x[: int(doc_tok[c])].cpu().numpy()
                first = [host[int(doc_tok[i]):int(doc_tok[i + 1])] for i in range(c)]
        else:
            assert nifs.mv_insert_device(idx, ids, x.data_ptr(), args.tokens, args.dim) == ("ok", ())
            if first is None:
                first = list(x[: args.check * args.tokens].cpu().numpy().reshape(-1, args.tokens, args.dim))
        del x
    torch.cuda.synchronize()
    ingest = time.perf_counter() - t0
    q = torch.randn(args.tq, args.dim, generator=g, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).cpu().numpy()

    st, hits = nifs.mv_search(idx, q, args.k)
    assert st == "ok", hits
    # parity spot check on the first `check` docs
    import oracle
    docs = [(f"{i:09d}", first[i]) for i in range(len(first))]
    ref = oracle.multi_vector_top_k(docs, q, nifs.METRIC_CODE[args.metric], args.k)[1]
    ref_d = dict(ref)
    for hid, s in hits:
        if int(hid) < len(first):
            assert hid in ref_d and abs(ref_d[hid] - s) <= 1e-5 * max(1.0, abs(s)), (hid, s, ref_d.get(hid))

    for _ in range(3):
        nifs.mv_search(idx, q, args.k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nifs.mv_search(idx, q, args.k)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    alg = total_tokens * args.dim * 4
    path = {0: "general", 1: "tcgen05 3xTF32 (uniform)", 2: "tcgen05 3xTF32 (ragged)"}.get(_lib.lib().vb_debug_maxsim_path(), "?")
    peak = 6545.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flops = 2.0 * args.tq * total_tokens * args.dim
    print(json.dumps({"metric": "maxsim queries/s (e2e through vb_mv_search)", "value": 1.0 / dt, "ms_per_query": dt * 1e3,
                      "config": {"docs": args.docs, "tokens": args.tokens, "dim": args.dim, "tq": args.tq, "k": args.k,
                                 "metric": args.metric, "ragged": args.ragged, "tokens_total": total_tokens, "kernel": path},
                      "roofline": {"bound": "hbm", "achieved": alg / dt / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / dt / 1e9 / peak, "algorithmic_tflops": flops / dt / 1e12},
                      "ingest_seconds": round(ingest, 2), "top1": hits[0]}))


if __name__ == "__main__":
    main()
