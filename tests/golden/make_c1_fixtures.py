#!/usr/bin/env python
"""Writes tests/golden/c1_fixtures.json: answers of the CPU oracle on seeded inputs of BASELINE.json's
CPU-runnable configuration (configs[0]: 10k x 384 fp32, L2-normalised, single query, k = 10) and on small
sign-code / MaxSim inputs. The inputs are NOT stored: `inputs()` below rebuilds them from the seed (numpy
Generator, PCG64), and tests/test_golden_fixtures.py imports it. The fixture pins the oracle against silent
regressions (CPU test) and is what the CUDA path is compared with on the GPU box (where /root/reference and a
second implementation are not available). Run: python tests/golden/make_c1_fixtures.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20_260_721   # the reference's own bench seed (bench/search_modes_bench.exs:14)


def unit(x):
    return (x / np.linalg.norm(x.astype(np.float64), axis=-1, keepdims=True)).astype(np.float32)


def inputs():
    rng = np.random.default_rng(SEED)
    rows = unit(rng.standard_normal((10_000, 384)).astype(np.float32))
    query = unit(rng.standard_normal(384).astype(np.float32))
    ids = [f"{i:09d}" for i in range(10_000)]
    docs = unit(rng.standard_normal((200, 8, 32)).astype(np.float32))
    mv_query = unit(rng.standard_normal((4, 32)).astype(np.float32))
    doc_ids = [f"doc-{i:04d}" for i in range(200)]
    return {"rows": rows, "query": query, "ids": ids, "docs": docs, "mv_query": mv_query, "doc_ids": doc_ids}


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle
    d = inputs()
    out = {"seed": SEED, "flat": {}, "binary": None, "quantized": None, "maxsim": {}}
    for metric in ("cosine", "l2", "inner_product", "manhattan"):
        st, hits = oracle.flat_search_dense(metric, d["rows"], d["ids"], d["query"], 10)
        assert st == "ok"
        out["flat"][metric] = [[i, float(v)] for i, v in hits]
    codes = [(d["ids"][i], oracle.compress_sign_bits(d["rows"][i])) for i in range(len(d["ids"]))]
    qcode = oracle.compress_sign_bits(d["query"])
    st, cand = oracle.binary_top_k(codes, qcode, 384, 100)
    assert st == "ok"
    out["binary"] = [[i, float(v)] for i, v in cand]
    keep = {i for i, _ in cand}
    st, top = oracle.vector_top_k([(d["ids"][i], d["rows"][i]) for i in range(len(d["ids"])) if d["ids"][i] in keep], d["query"], 2, 384, 10)
    assert st == "ok"
    out["quantized"] = [[i, float(v)] for i, v in top]
    for code, name in ((3, "inner_product"), (2, "cosine"), (0, "l2")):
        st, hits = oracle.multi_vector_top_k([(d["doc_ids"][i], d["docs"][i]) for i in range(200)], d["mv_query"], code, 5)
        assert st == "ok"
        out["maxsim"][name] = [[i, float(v)] for i, v in hits]
    path = os.path.join(HERE, "c1_fixtures.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
