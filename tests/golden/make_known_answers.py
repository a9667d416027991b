#!/usr/bin/env python
"""Writes tests/golden/reference_known_answers.json: the known-answer cases of the reference's OWN tests for
the hot path, transcribed by hand (the reference is Rust + Elixir and cannot run in this image: no cargo, no
erl). Every case cites the reference test it comes from (paths under native/vettore/src/ or test/).

Case format: {"fn": name, "args": [...], "expect": {"ok": value} | {"error": message} | {"ok_ids": [...]},
"source": "file:lines"}. `fn` names exist with the same signature in `oracle` (CPU restatement) and in
`vettore_b200.nifs` (CUDA path through the C ABI); tests/test_golden_fixtures.py runs every case through both.
Hits are [[id, value], ...]. "flat_script" cases are a list of index operations.
Run: python tests/golden/make_known_answers.py"""
import json
import os

M64 = (1 << 64) - 1
NAN, INF = "nan", "inf"     # spelled as strings in JSON, decoded by the runner

cases = [
    # ---- search.rs: vector_top_k
    {"fn": "vector_top_k", "args": [[["b", [1.0, 10.0]], ["a", [1.0, -10.0]], ["c", [-1.0, 0.0]]], [1.0, 0.0], 0, 1, 2],
     "expect": {"ok": [["a", 0.0], ["b", 0.0]]}, "source": "search.rs:158-173 (prefix of 1 dimension, tie broken by id)"},
    {"fn": "vector_top_k", "args": [[["b", [1.0, 10.0]], ["a", [1.0, -10.0]], ["c", [-1.0, 0.0]]], [1.0, 1.0], 3, 2, 1],
     "expect": {"ok": [["b", 11.0]]}, "source": "search.rs:158-173 (inner product ranks the largest dot first)"},
    {"fn": "vector_top_k", "args": [[], [1.0], 0, 0, 1], "expect": {"error": "invalid prefix dimensions"}, "source": "search.rs:175-184"},
    {"fn": "vector_top_k", "args": [[["a", [1.0]]], [1.0, 2.0], 0, 2, 1], "expect": {"error": "dimension mismatch"}, "source": "search.rs:175-184"},
    {"fn": "vector_top_k", "args": [[["a", [NAN]]], [1.0], 0, 1, 1], "expect": {"error": "vector contains a non-finite value"},
     "source": "search.rs:175-184"},
    {"fn": "vector_top_k", "args": [[["b", [1.0, 0.0]], ["a", [1.0, 0.0]], ["c", [0.0, 1.0]]], [1.0, 0.0], 9, 2, 2],
     "expect": {"error": "unknown metric"}, "source": "test/vector_algorithms_hardening_test.exs:90-121"},
    {"fn": "vector_top_k", "args": [[["b", [1.0, 0.0]], ["a", [1.0, 0.0]], ["c", [0.0, 1.0]]], [1.0, 0.0], 0, 0, 2],
     "expect": {"error": "invalid prefix dimensions"}, "source": "test/vector_algorithms_hardening_test.exs:90-121"},
    # ---- search.rs: binary_top_k, distances.rs: packing
    {"fn": "binary_top_k", "args": [[["b", [1]], ["a", [3]]], [3], 2, 2], "expect": {"ok": [["a", 0.0], ["b", 1.0]]},
     "source": "test/vector_algorithms_hardening_test.exs:90-121"},
    {"fn": "binary_top_k", "args": [[["b", [7]], ["a", [5]]], [5], 3, 2], "expect": {"ok": [["a", 0.0], ["b", 1.0]]},
     "source": "search.rs:186-203 (sign codes of [1,1,1] / [1,-1,1] against [1,-1,1])"},
    {"fn": "binary_top_k", "args": [[], [], 0, 1], "expect": {"error": "dimensions must be positive"}, "source": "search.rs:246-260"},
    {"fn": "binary_top_k", "args": [[], [], 1, 1], "expect": {"error": "dimension mismatch"}, "source": "search.rs:246-260"},
    {"fn": "binary_top_k", "args": [[], [0], 1, 1], "expect": {"ok": []}, "source": "search.rs:246-260"},
    {"fn": "binary_top_k", "args": [[["same", [M64, 1]], ["far", [0, 0]]], [M64, 1], 65, 0], "expect": {"ok": []}, "source": "search.rs:246-260"},
    {"fn": "binary_top_k", "args": [[["same", [M64, 1]], ["far", [0, 0]]], [M64, 1], 65, 10],
     "expect": {"ok": [["same", 0.0], ["far", 65.0]]}, "source": "search.rs:246-260 (65 bits: word boundary + last-word mask)"},
    {"fn": "binary_top_k", "args": [[["bad", [0]]], [M64, 1], 65, 1], "expect": {"error": "dimension mismatch"}, "source": "search.rs:246-260"},
    {"fn": "compress_sign_bits", "args": [[1.0, -1.0, 0.0]], "expect": {"ok": [5]}, "source": "distances.rs:539-548; vettore_distance.ex:253-254"},
    {"fn": "compress_sign_bits", "args": [[-1.0, -1.0, 0.0]], "expect": {"ok": [4]}, "source": "distances.rs:539-548"},
    {"fn": "compress_sign_bits", "args": [[-0.0]], "expect": {"ok": [1]}, "source": "distances.rs:413-423 (`>= 0.0` holds for -0.0)"},
    # ---- multi_vector.rs
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]], 3], "expect": {"ok": 2.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]], 4], "expect": {"ok": 2.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]], 2], "expect": {"ok": 2.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]], 0], "expect": {"ok": 2.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_score", "args": [[], [[1.0, 0.0], [0.0, 1.0]], 0], "expect": {"ok": 0.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0], [0.0, 1.0]], [], 0], "expect": {"ok": 0.0}, "source": "multi_vector.rs:193-206"},
    {"fn": "multi_vector_top_k", "args": [[["b", [[1.0, 0.0]]], ["a", [[1.0, 0.0]]], ["c", [[-1.0, 0.0]]]], [[1.0, 0.0]], 3, 2],
     "expect": {"ok": [["a", 1.0], ["b", 1.0]]}, "source": "multi_vector.rs:208-222 (score ties resolve to the smaller id)"},
    {"fn": "multi_vector_score", "args": [[[1.0, 0.0]], [[1.0]], 3], "expect": {"error": "dimension mismatch"}, "source": "multi_vector.rs:208-222"},
    {"fn": "multi_vector_score", "args": [[[NAN, 0.0]], [[1.0, 0.0]], 3], "expect": {"error": "vector contains a non-finite value"},
     "source": "multi_vector.rs:208-222"},
    # ---- flat.rs: index scripts
    {"fn": "flat_script", "args": ["l2", [["insert", "b", [2.0]], ["insert", "a", [0.0]], ["insert", "c", [2.0]],
                                          ["search", [1.0], 2, {"ok": [["a", 1.0], ["b", 1.0]]}],
                                          ["insert", "a", [10.0]], ["search", [2.0], 1, {"ok_ids": ["b"]}],
                                          ["delete", "b"], ["search", [2.0], 1, {"ok_ids": ["c"]}]]],
     "expect": {"ok": None}, "source": "flat.rs:164-180 (upsert, delete, stable top-k)"},
    {"fn": "flat_script", "args": ["inner_product", [["insert", "existing", [1.0, 0.0]],
                                                     ["insert_many", [["valid", [0.0, 1.0]], ["invalid", [1.0]]], {"error": "dimension mismatch"}],
                                                     ["search", [1.0, 0.0], 5, {"ok_ids": ["existing"]}],
                                                     ["insert", "not-finite", [NAN, 0.0], {"error": "vector contains a non-finite value"}]]],
     "expect": {"ok": None}, "source": "flat.rs:182-196 (batch validation is all-or-nothing)"},
    {"fn": "flat_script", "args": ["cosine", [["insert", "empty", [], {"error": "vector must not be empty"}],
                                              ["insert", "a", [1.0, 0.0]],
                                              ["search", [1.0], 1, {"error": "dimension mismatch"}],
                                              ["search", [INF, 0.0], 1, {"error": "vector contains a non-finite value"}],
                                              ["search", [1.0, 0.0], 0, {"ok": []}]]],
     "expect": {"ok": None}, "source": "flat.rs:198-206"},
]
for code in range(9):
    cases.append({"fn": "vector_top_k", "args": [[["b", [1.0, 0.0]], ["a", [1.0, 0.0]], ["c", [0.0, 1.0]]], [1.0, 0.0], code, 2, 2],
                  "expect": {"ok_ids": ["a", "b"]}, "source": "test/vector_algorithms_hardening_test.exs:90-121 (every metric code, ties by id)"})
for metric in ["l2", "l2_squared", "cosine", "inner_product", "negative_inner_product", "manhattan", "chebyshev", "hamming", "jaccard"]:
    cases.append({"fn": "flat_script", "args": [metric, [["insert_many", [["b", [0.0, 1.0]], ["a", [1.0, 0.0]], ["c", [1.0, 0.0]]]],
                                                          ["search", [1.0, 0.0], 2, {"ok_ids": ["a", "c"]}]]],
                  "expect": {"ok": None}, "source": "test/vector_algorithms_hardening_test.exs:20-36 (stable ties for every flat metric)"})

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
with open(out, "w") as f:
    json.dump({"note": "transcribed from the reference's own tests; see make_known_answers.py", "cases": cases}, f, indent=1)
print(len(cases), "cases ->", out)
