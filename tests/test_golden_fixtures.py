"""Runs the committed golden fixtures (tests/golden/) through BOTH implementations:
  * the CPU oracle (always; pins the restatement to the reference's known answers and to its own seeded answers)
  * the CUDA path through the C ABI (`-m gpu`; the same cases, the same expectations).
reference_known_answers.json is transcribed from the reference's own tests (make_known_answers.py);
c1_fixtures.json holds oracle answers on BASELINE.json configs[0] (make_c1_fixtures.py rebuilds the inputs)."""
import importlib.util
import json
import os

import numpy as np
import pytest

import oracle
from helpers import assert_hits_match, close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))["cases"]
C1 = json.load(open(os.path.join(GOLDEN, "c1_fixtures.json")))


def _decode(x):
    """JSON spells non-finite floats as strings."""
    if isinstance(x, str) and x in ("nan", "inf", "-inf"):
        return float(x)
    if isinstance(x, list):
        return [_decode(v) for v in x]
    return x


def _pairs(rows):
    return [(r[0], _decode(r[1])) for r in rows]


def _check(result, expect, what):
    if "error" in expect:
        assert result == ("error", expect["error"]), (what, result)
        return
    assert result[0] == "ok", (what, result)
    if "ok_ids" in expect:
        assert [h[0] for h in result[1]] == expect["ok_ids"], (what, result)
    elif expect["ok"] is None:
        return
    elif isinstance(expect["ok"], list) and expect["ok"] and isinstance(expect["ok"][0], list):
        got = [(h[0], float(h[1])) for h in result[1]]
        assert got == [(i, float(v)) for i, v in expect["ok"]], (what, got)
    elif isinstance(expect["ok"], list):
        assert list(result[1]) == expect["ok"], (what, result)
    else:
        assert float(result[1]) == float(expect["ok"]), (what, result)


class _OracleFlat:
    def __init__(self, metric):
        self.idx = oracle.FlatIndex(metric)
    def insert(self, i, v): return self.idx.insert(i, v)
    def insert_many(self, items): return self.idx.insert_many(items)
    def delete(self, i): return self.idx.delete(i)
    def search(self, q, k): return self.idx.search(q, k)


class _CudaFlat:
    def __init__(self, metric):
        from vettore_b200 import nifs
        self.n, self.idx = nifs, getattr(nifs, f"flat_new_{metric}")()
    def insert(self, i, v): return self.n.flat_insert(self.idx, i, v)
    def insert_many(self, items): return self.n.flat_insert_many(self.idx, items)
    def delete(self, i): return self.n.flat_delete(self.idx, i)
    def search(self, q, k): return self.n.flat_search(self.idx, q, k)


def _run_case(case, impl, flat_cls):
    fn, args, what = case["fn"], _decode(case["args"]), case["source"]
    if fn == "flat_script":
        metric, ops = args
        idx = flat_cls(metric)
        for op in ops:
            if op[0] == "insert":
                res, exp = idx.insert(op[1], op[2]), (op[3] if len(op) > 3 else {"ok": None})
            elif op[0] == "insert_many":
                res, exp = idx.insert_many(_pairs(op[1])), (op[2] if len(op) > 2 else {"ok": None})
            elif op[0] == "delete":
                res, exp = idx.delete(op[1]), {"ok": None}
                res = ("ok", None) if res is None or res[0] == "ok" else res
            else:
                res, exp = idx.search(op[1], op[2]), op[3]
            _check(res, exp, (what, op[0]))
        return
    f = getattr(impl, fn)
    if fn in ("vector_top_k", "binary_top_k", "multi_vector_top_k"):
        args = [_pairs(args[0])] + list(args[1:])
    res = f(*args)
    if fn == "compress_sign_bits":
        res = ("ok", [int(w) for w in res])
    _check(res, case["expect"], what)


@pytest.mark.parametrize("case", KNOWN, ids=lambda c: f"{c['fn']}@{c['source'].split(' ')[0]}")
def test_oracle_reproduces_the_reference_known_answers(case):
    _run_case(case, oracle, _OracleFlat)


@pytest.mark.gpu
@pytest.mark.parametrize("case", KNOWN, ids=lambda c: f"{c['fn']}@{c['source'].split(' ')[0]}")
def test_cuda_path_reproduces_the_reference_known_answers(case):
    from vettore_b200 import nifs
    _run_case(case, nifs, _CudaFlat)


def _c1_inputs():
    spec = importlib.util.spec_from_file_location("make_c1_fixtures", os.path.join(GOLDEN, "make_c1_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.inputs()


def _hits(rows):
    return [(i, float(v)) for i, v in rows]


def test_oracle_matches_its_committed_c1_answers():
    d = _c1_inputs()
    for metric, want in C1["flat"].items():
        st, hits = oracle.flat_search_dense(metric, d["rows"], d["ids"], d["query"], 10)
        assert st == "ok" and [(i, float(v)) for i, v in hits] == _hits(want), metric
    codes = [(d["ids"][i], oracle.compress_sign_bits(d["rows"][i])) for i in range(len(d["ids"]))]
    st, cand = oracle.binary_top_k(codes, oracle.compress_sign_bits(d["query"]), 384, 100)
    assert st == "ok" and [(i, float(v)) for i, v in cand] == _hits(C1["binary"])


@pytest.mark.gpu
def test_cuda_path_matches_the_committed_c1_answers():
    """BASELINE.json configs[0] (10k x 384, one query, k = 10) and the small quantized / MaxSim fixtures:
    float scores within 1e-5, Hamming candidates bit-exact."""
    from vettore_b200 import nifs
    d = _c1_inputs()
    for metric, want in C1["flat"].items():
        idx = getattr(nifs, f"flat_new_{metric}")()
        assert nifs.flat_insert_matrix(idx, d["ids"], d["rows"]) == ("ok", ())
        st, hits = nifs.flat_search(idx, d["query"], 10)
        assert st == "ok"
        assert_hits_match(hits, _hits(want))
        if metric == "cosine":
            st, top = nifs.flat_quantized_search(idx, d["query"], nifs.METRIC_CODE["cosine"], 100, 10)
            assert st == "ok"
            assert_hits_match(top, _hits(C1["quantized"]))
    codes = [(d["ids"][i], nifs.compress_sign_bits(d["rows"][i])) for i in range(len(d["ids"]))]
    st, cand = nifs.binary_top_k(codes, nifs.compress_sign_bits(d["query"]), 384, 100)
    assert st == "ok" and [(i, float(v)) for i, v in cand] == _hits(C1["binary"])      # integers: exact
    for name, want in C1["maxsim"].items():
        st, hits = nifs.multi_vector_top_k([(d["doc_ids"][i], d["docs"][i]) for i in range(200)], d["mv_query"],
                                           nifs.METRIC_CODE[name], 5)
        assert st == "ok"
        assert_hits_match(hits, _hits(want))
        mv = nifs.mv_new(name)
        assert nifs.mv_insert_tensor(mv, d["doc_ids"], d["docs"])[0] == "ok"
        st, hits = nifs.mv_search(mv, d["mv_query"], 5)
        assert st == "ok"
        assert_hits_match(hits, _hits(want))
