"""CPU oracle for Vettore's scan path — TEST INFRASTRUCTURE, never the product path.

Thin ctypes binding over ``oracle/liboracle.so`` (built from ``vettore_oracle.cpp``,
a C++ restatement of ``native/vettore/src/{distances,flat,search,multi_vector}.rs``)
plus the small stateful pieces of ``flat.rs:48-93`` (upsert / delete / dimension
tracking) that are plain dictionary logic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package. Results use the reference's NIF conventions: ``("ok", value)`` or
``("error", "message")`` with the reference's exact error strings.

Pinning: the Rust reference cannot be built here (no cargo/rustc/erl). The oracle is
pinned against the reference's own known-answer tests, restated in
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

METRICS = ["l2", "l2_squared", "cosine", "inner_product", "negative_inner_product",
           "manhattan", "chebyshev", "hamming", "jaccard"]  # distances.rs:25-38
METRIC_CODE = {m: i for i, m in enumerate(METRICS)}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vettore_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vo_last_error.restype = C.c_char_p
        L.vo_rank_value.restype = C.c_float
        L.vo_rank_value.argtypes = [C.c_uint8, C.c_float]
        L.vo_similarity_value.restype = C.c_float
        L.vo_similarity_value.argtypes = [C.c_uint8, C.c_float]
        L.vo_flat_scan_timed.restype = C.c_double
        _lib = L
    return _lib


def _err():
    return ("error", lib().vo_last_error().decode())


def _f32(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))


def _u64(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray([int(x) for x in v], dtype=np.uint64).reshape(-1))


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _ids_blob(ids: Sequence[str | bytes]):
    enc = [i.encode() if isinstance(i, str) else bytes(i) for i in ids]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    blob = b"".join(enc)
    buf = C.create_string_buffer(blob, max(len(blob), 1))
    return buf, off


def _ragged_f32(rows: Iterable[Sequence[float]]):
    rows = [np.asarray(r, dtype=np.float32).reshape(-1) for r in rows]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    if rows:
        off[1:] = np.cumsum([len(r) for r in rows], dtype=np.uint64)
    vals = np.concatenate(rows) if rows else np.zeros(0, dtype=np.float32)
    return np.ascontiguousarray(vals, dtype=np.float32), off


def _ragged_u64(rows: Iterable[Sequence[int]]):
    rows = [np.asarray([int(x) for x in r], dtype=np.uint64).reshape(-1) for r in rows]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    if rows:
        off[1:] = np.cumsum([len(r) for r in rows], dtype=np.uint64)
    vals = np.concatenate(rows) if rows else np.zeros(0, dtype=np.uint64)
    return np.ascontiguousarray(vals, dtype=np.uint64), off


# --------------------------------------------------------------------------- pairwise
def compute(metric: str | int, left, right, checked: bool = False):
    """distances.rs:42-68 (``checked`` = compute_checked, :101-105)."""
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    if not 0 <= code <= 8:
        return ("error", "unknown metric")
    a, b = _f32(left), _f32(right)
    out = C.c_float()
    rc = lib().vo_compute(C.c_uint8(code), _p(a, C.c_float), C.c_size_t(a.size), _p(b, C.c_float),
                          C.c_size_t(b.size), C.c_int(int(checked)), C.byref(out))
    return _err() if rc else ("ok", out.value)


def cosine(left, right):
    """distances.rs:160-177 (true cosine in f64)."""
    a, b = _f32(left), _f32(right)
    out = C.c_float()
    rc = lib().vo_cosine(_p(a, C.c_float), C.c_size_t(a.size), _p(b, C.c_float), C.c_size_t(b.size), C.byref(out))
    return _err() if rc else ("ok", out.value)


def rank_value(metric: str | int, raw: float) -> float:
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    return lib().vo_rank_value(code, raw)


def similarity_value(metric: str | int, raw: float) -> float:
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    return lib().vo_similarity_value(code, raw)


def normalize_l2(v):
    a = _f32(v)
    out = np.zeros_like(a)
    rc = lib().vo_normalize_l2(_p(a, C.c_float), C.c_size_t(a.size), _p(out, C.c_float))
    return _err() if rc else ("ok", out)


def compress_sign_bits(v) -> list[int]:
    """distances.rs:413-423."""
    a = _f32(v)
    words = np.zeros((a.size + 63) // 64, dtype=np.uint64)
    lib().vo_compress_sign_bits(_p(a, C.c_float), C.c_size_t(a.size), _p(words, C.c_uint64))
    return [int(w) for w in words]


def packed_hamming(left, right, dimensions: int):
    a, b = _u64(left), _u64(right)
    out = C.c_float()
    rc = lib().vo_packed_hamming(_p(a, C.c_uint64), C.c_size_t(a.size), _p(b, C.c_uint64), C.c_size_t(b.size),
                                 C.c_size_t(dimensions), C.byref(out))
    return _err() if rc else ("ok", out.value)


def packed_jaccard(left, right, dimensions: int):
    a, b = _u64(left), _u64(right)
    out = C.c_float()
    rc = lib().vo_packed_jaccard(_p(a, C.c_uint64), C.c_size_t(a.size), _p(b, C.c_uint64), C.c_size_t(b.size),
                                 C.c_size_t(dimensions), C.byref(out))
    return _err() if rc else ("ok", out.value)


# --------------------------------------------------------------------------- batched
def _hits(ids, idx, raw, n):
    return [(ids[int(idx[i])], float(raw[i])) for i in range(n)]


def vector_top_k(vectors: Sequence[tuple[str, Sequence[float]]], query, metric_code: int, dimensions: int, limit: int):
    """search.rs:38-73 behind nifs.rs:151-162."""
    ids = [v[0] for v in vectors]
    vals, off = _ragged_f32([v[1] for v in vectors])
    blob, ioff = _ids_blob(ids)
    q = _f32(query)
    cap = max(1, min(limit, len(ids)))
    oi, orr, on = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_size_t()
    rc = lib().vo_vector_top_k(_p(vals, C.c_float), _p(off, C.c_uint64), C.c_size_t(len(ids)), blob, _p(ioff, C.c_uint64),
                               _p(q, C.c_float), C.c_size_t(q.size), C.c_int(metric_code), C.c_size_t(dimensions),
                               C.c_size_t(limit), _p(oi, C.c_uint64), _p(orr, C.c_float), C.byref(on))
    return _err() if rc else ("ok", _hits(ids, oi, orr, on.value))


def binary_top_k(vectors: Sequence[tuple[str, Sequence[int]]], query, dimensions: int, limit: int):
    """search.rs:76-92 behind nifs.rs:164-175."""
    ids = [v[0] for v in vectors]
    vals, off = _ragged_u64([v[1] for v in vectors])
    blob, ioff = _ids_blob(ids)
    q = _u64(query)
    cap = max(1, min(limit, len(ids)))
    oi, orr, on = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_size_t()
    rc = lib().vo_binary_top_k(_p(vals, C.c_uint64), _p(off, C.c_uint64), C.c_size_t(len(ids)), blob, _p(ioff, C.c_uint64),
                               _p(q, C.c_uint64), C.c_size_t(q.size), C.c_size_t(dimensions), C.c_size_t(limit),
                               _p(oi, C.c_uint64), _p(orr, C.c_float), C.byref(on))
    return _err() if rc else ("ok", _hits(ids, oi, orr, on.value))


def multi_vector_score(query_vectors, document_vectors, metric_code: int):
    """multi_vector.rs:40-63 behind nifs.rs:177-186."""
    qv, qoff = _ragged_f32(query_vectors)
    dv, doff = _ragged_f32(document_vectors)
    out = C.c_float()
    rc = lib().vo_multi_vector_score(_p(qv, C.c_float), _p(qoff, C.c_uint64), C.c_size_t(len(qoff) - 1),
                                     _p(dv, C.c_float), _p(doff, C.c_uint64), C.c_size_t(len(doff) - 1),
                                     C.c_int(metric_code), C.byref(out))
    return _err() if rc else ("ok", out.value)


def multi_vector_top_k(documents: Sequence[tuple[str, Sequence[Sequence[float]]]], query_vectors, metric_code: int, limit: int):
    """multi_vector.rs:90-132 behind nifs.rs:188-198."""
    ids = [d[0] for d in documents]
    toks, doc_tok = [], np.zeros(len(ids) + 1, dtype=np.uint64)
    for i, (_, vs) in enumerate(documents):
        toks.extend(vs)
        doc_tok[i + 1] = len(toks)
    dv, doff = _ragged_f32(toks)
    qv, qoff = _ragged_f32(query_vectors)
    blob, ioff = _ids_blob(ids)
    cap = max(1, min(limit, len(ids)))
    oi, orr, on = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_size_t()
    rc = lib().vo_multi_vector_top_k(_p(dv, C.c_float), _p(doff, C.c_uint64), _p(doc_tok, C.c_uint64), C.c_size_t(len(ids)),
                                     blob, _p(ioff, C.c_uint64), _p(qv, C.c_float), _p(qoff, C.c_uint64),
                                     C.c_size_t(len(qoff) - 1), C.c_int(metric_code), C.c_size_t(limit),
                                     _p(oi, C.c_uint64), _p(orr, C.c_float), C.byref(on))
    return _err() if rc else ("ok", _hits(ids, oi, orr, on.value))


# --------------------------------------------------------------------------- flat index
class FlatIndex:
    """Restates ``FlatIndex`` (flat.rs:13-129): an id -> vector map, one metric, one dimension."""

    def __init__(self, metric: str | int):
        self.metric = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
        self.vectors: dict[str, np.ndarray] = {}
        self.dimension: int | None = None

    @staticmethod
    def _validate(vector: np.ndarray, dimension: int | None):  # flat.rs:136-144
        if vector.size == 0:
            return "vector must not be empty"
        if dimension is not None and vector.size != dimension:
            return "dimension mismatch"
        if not np.all(np.isfinite(vector)):
            return "vector contains a non-finite value"
        return None

    def insert(self, id: str, vector):  # flat.rs:59-66
        v = _f32(vector)
        e = self._validate(v, self.dimension)
        if e:
            return ("error", e)
        if self.dimension is None:
            self.dimension = v.size
        self.vectors[id] = v
        return ("ok", ())

    def insert_many(self, vectors: Sequence[tuple[str, Sequence[float]]]):  # flat.rs:69-85
        vs = [(i, _f32(v)) for i, v in vectors]
        expected = self.dimension if self.dimension is not None else (vs[0][1].size if vs else None)
        for _, v in vs:
            e = self._validate(v, expected)
            if e:
                return ("error", e)
        for i, v in vs:
            self.vectors[i] = v
        if self.dimension is None:
            self.dimension = expected
        return ("ok", ())

    def delete(self, id: str):  # flat.rs:88-93
        self.vectors.pop(id, None)
        if not self.vectors:
            self.dimension = None
        return ("ok", ())

    def search(self, query, limit: int):  # flat.rs:96-124
        ids = list(self.vectors.keys())
        q = _f32(query)
        d = self.dimension or 0
        rows = np.ascontiguousarray(np.stack([self.vectors[i] for i in ids]) if ids else np.zeros((0, max(d, 1)), np.float32))
        return flat_search_dense(self.metric, rows, ids, q, limit, index_dim=self.dimension)


def flat_search_dense(metric: str | int, rows: np.ndarray, ids: Sequence[str] | None, query, limit: int,
                      index_dim: int | None = -2):
    """flat.rs:96-124 over a dense [n, d] snapshot. ``ids=None`` means zero-padded row numbers."""
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    if ids is None:
        width = max(9, len(str(max(n - 1, 0))))
        ids = [f"{i:0{width}d}" for i in range(n)]
    if index_dim == -2:
        index_dim = d if n else None
    blob, ioff = _ids_blob(ids)
    q = _f32(query)
    cap = max(1, min(limit, n))
    oi, orr, on = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_size_t()
    rc = lib().vo_flat_search(C.c_uint8(code), _p(rows, C.c_float), C.c_size_t(n), C.c_size_t(d),
                              C.c_longlong(-1 if index_dim is None else index_dim), blob, _p(ioff, C.c_uint64),
                              _p(q, C.c_float), C.c_size_t(q.size), C.c_size_t(limit),
                              _p(oi, C.c_uint64), _p(orr, C.c_float), C.byref(on))
    return _err() if rc else ("ok", _hits(ids, oi, orr, on.value))


def flat_scan_timed(metric: str | int, rows: np.ndarray, queries: np.ndarray, limit: int, threads: int):
    """Timed CPU baseline (see vo_flat_scan_timed). Returns (seconds, per query [(row, raw)]) — row order is
    id order for ids that sort like the row number (the bench's zero-padded decimals)."""
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n, d = rows.shape
    nq = queries.shape[0]
    cap = max(1, min(limit, n))
    oi, orr = np.zeros(nq * cap, np.uint64), np.zeros(nq * cap, np.float32)
    secs = lib().vo_flat_scan_timed(C.c_uint8(code), _p(rows, C.c_float), C.c_size_t(n), C.c_size_t(d),
                                    _p(queries, C.c_float), C.c_size_t(nq), C.c_size_t(limit), C.c_int(threads),
                                    _p(oi, C.c_uint64), _p(orr, C.c_float))
    return secs, [[(int(oi[q * cap + i]), float(orr[q * cap + i])) for i in range(cap)] for q in range(nq)]


def binary_scan_timed(codes: np.ndarray, dims: int, queries: np.ndarray, limit: int, threads: int):
    """Timed CPU baseline of the Hamming candidate pass (vo_binary_scan_timed; search.rs:76-92) over a dense
    ``[n, nw]`` uint64 code matrix. Returns (seconds, per query [(row, distance)])."""
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    queries = np.ascontiguousarray(queries, dtype=np.uint64)
    n, nw = codes.shape
    nq = queries.shape[0]
    cap = max(1, min(limit, n))
    oi, orr = np.zeros(nq * cap, np.uint64), np.zeros(nq * cap, np.float32)
    fn = lib().vo_binary_scan_timed
    fn.restype = C.c_double
    secs = fn(_p(codes, C.c_uint64), C.c_size_t(n), C.c_size_t(nw), C.c_size_t(dims), _p(queries, C.c_uint64),
              C.c_size_t(nq), C.c_size_t(limit), C.c_int(threads), _p(oi, C.c_uint64), _p(orr, C.c_float))
    return secs, [[(int(oi[q * cap + i]), float(orr[q * cap + i])) for i in range(cap)] for q in range(nq)]


def maxsim_scan_timed(metric: str | int, tokens: np.ndarray, queries: np.ndarray, limit: int, threads: int):
    """Timed CPU baseline of MaxSim top-k (vo_maxsim_scan_timed; multi_vector.rs:90-132) over uniform documents
    ``[ndocs, td, dim]`` and queries ``[nq, tq, dim]``. Returns (seconds, per query [(doc, score)])."""
    code = METRIC_CODE[metric] if isinstance(metric, str) else int(metric)
    tokens = np.ascontiguousarray(tokens, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    ndocs, td, dim = tokens.shape
    nq, tq, _ = queries.shape
    cap = max(1, min(limit, ndocs))
    oi, orr = np.zeros(nq * cap, np.uint64), np.zeros(nq * cap, np.float32)
    fn = lib().vo_maxsim_scan_timed
    fn.restype = C.c_double
    secs = fn(_p(tokens, C.c_float), C.c_size_t(ndocs), C.c_size_t(td), C.c_size_t(dim), _p(queries, C.c_float),
              C.c_size_t(nq), C.c_size_t(tq), C.c_int(code), C.c_size_t(limit), C.c_int(threads),
              _p(oi, C.c_uint64), _p(orr, C.c_float))
    if secs < 0:
        raise RuntimeError("oracle MaxSim scan failed: " + _err()[1])
    return secs, [[(int(oi[q * cap + i]), float(orr[q * cap + i])) for i in range(cap)] for q in range(nq)]


def muvera_encode(vectors, dimension: int, num_repetitions: int, num_simhash_projections: int, seed: int,
                  projection_dimension: int, final_projection_dimension: int | None, mode: str):
    """muvera.rs:26-74 (nifs.rs:430-476: muvera_encode_query / muvera_encode_document). ``mode`` is "query" (sum per
    partition) or "document" (running average). Returns ("ok", [float]) or ("error", msg)."""
    vals, off = _ragged_f32(vectors)
    part = 1 << min(int(num_simhash_projections), 40)
    cap = int(final_projection_dimension) if final_projection_dimension else min(int(num_repetitions) * part *
                                                                                max(1, int(projection_dimension)), 16777216)
    out = np.zeros(max(1, cap), np.float32)
    n = C.c_size_t()
    rc = lib().vo_muvera_encode(_p(vals, C.c_float), _p(off, C.c_uint64), C.c_size_t(len(off) - 1), C.c_size_t(dimension),
                                C.c_size_t(num_repetitions), C.c_size_t(num_simhash_projections), C.c_uint64(seed),
                                C.c_size_t(projection_dimension), C.c_int(final_projection_dimension is not None),
                                C.c_size_t(final_projection_dimension or 0), C.c_int({"query": 0, "document": 1}[mode]),
                                _p(out, C.c_float), C.c_size_t(out.size), C.byref(n))
    return _err() if rc else ("ok", out[: n.value].tolist())
