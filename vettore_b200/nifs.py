"""Function-for-function mirror of ``Vettore.Nifs`` (reference lib/vettore_nifs.ex:16-258)
for the scan path, bound to the CUDA library through the C ABI.

Same names, argument order and result shapes as the reference NIFs:
``{:ok, value}`` -> ``("ok", value)``, ``{:error, msg}`` -> ``("error", msg)`` with the
reference's exact strings; ``flat_new_*`` return the bare resource like nifs.rs:200-257;
``Result<(), String>`` successes are ``("ok", ())`` like ``{:ok, {}}``.
Ids are ``str`` (UTF-8 binaries on the BEAM) or ``bytes``.

HNSW, MUVERA, the pairwise metric helpers and the normalisers stay on the reference
path (out of scope) and are not defined here.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import SIZE_MAX, lib

METRICS = ["l2", "l2_squared", "cosine", "inner_product", "negative_inner_product",
           "manhattan", "chebyshev", "hamming", "jaccard"]  # distances.rs:25-38
METRIC_CODE = {m: i for i, m in enumerate(METRICS)}

_f32p, _u64p, _u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)


def _err():
    return ("error", _lib.last_error())


def _f32(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _enc(i) -> bytes:
    return i.encode("utf-8") if isinstance(i, str) else bytes(i)


class IdBlob:
    """Ids already laid out the way the C ABI takes them (one byte blob + n+1 offsets): bulk loads of
    millions of rows skip the per-id Python objects. ``len()`` = number of ids."""

    def __init__(self, blob: np.ndarray, off: np.ndarray):
        self.blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.uint64)

    def __len__(self) -> int:
        return int(self.off.size) - 1

    def __getitem__(self, i: int) -> str:
        return bytes(self.blob[int(self.off[i]):int(self.off[i + 1])]).decode("utf-8")


def decimal_ids(base: int, n: int, width: int = 9) -> IdBlob:
    """Zero-padded decimal ids ``base .. base+n-1`` (id byte order == numeric order: SURVEY.md §8(d))."""
    blob = np.empty((n, width), dtype=np.uint8)
    chunk = 1 << 22
    pw = (10 ** np.arange(width - 1, -1, -1)).astype(np.int64)
    for s in range(0, n, chunk):
        v = np.arange(base + s, base + min(n, s + chunk), dtype=np.int64)
        blob[s:s + v.size] = ((v[:, None] // pw[None, :]) % 10 + 48).astype(np.uint8)
    return IdBlob(blob.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(width))


def _ids_blob(ids) -> tuple:
    if isinstance(ids, IdBlob):
        return ids.blob.ctypes.data_as(C.c_char_p), ids.off
    enc = [_enc(i) for i in ids]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        off[1:] = np.cumsum(np.fromiter((len(e) for e in enc), dtype=np.uint64, count=len(enc)))
    return b"".join(enc), off


def _ragged(rows: Iterable, dtype) -> tuple[np.ndarray, np.ndarray]:
    if isinstance(rows, np.ndarray) and rows.ndim == 2:
        n, d = rows.shape
        vals = np.ascontiguousarray(rows, dtype=dtype).reshape(-1)
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(d)
        return vals, off
    if dtype == np.uint64:
        rows = [np.asarray([int(x) for x in r], dtype=np.uint64).reshape(-1) for r in rows]
    else:
        rows = [np.asarray(r, dtype=dtype).reshape(-1) for r in rows]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    if rows:
        off[1:] = np.cumsum(np.fromiter((r.size for r in rows), dtype=np.uint64, count=len(rows)))
    vals = np.concatenate(rows) if rows else np.zeros(0, dtype=dtype)
    return np.ascontiguousarray(vals, dtype=dtype), off


def _take_hits(handle: C.c_void_p, as_str: bool = True) -> list[tuple]:
    L = lib()
    try:
        blob, off, vals, idx = C.c_void_p(), _u64p(), _f32p(), _u64p()
        n = L.vb_hits_export(handle, C.byref(blob), C.byref(off), C.byref(vals), C.byref(idx))
        if n == 0:
            return []
        offs = off[: n + 1]
        raw = C.string_at(blob.value, offs[n])
        if as_str:
            return [(raw[offs[i]:offs[i + 1]].decode("utf-8"), vals[i]) for i in range(n)]
        return [(raw[offs[i]:offs[i + 1]], vals[i]) for i in range(n)]
    finally:
        L.vb_hits_free(handle)


class FlatRef:
    """Opaque resource handle (the BEAM ``reference()`` of flat_new_*). Freed on GC."""

    def __init__(self, handle: C.c_void_p, metric: str):
        self._h = handle
        self.metric = metric

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().vb_flat_free(h)
            except Exception:
                pass

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise ValueError("flat index resource already released")
        return self._h


def _flat_new(metric: str) -> FlatRef:
    h = C.c_void_p()
    rc = lib().vb_flat_new(METRIC_CODE[metric], C.byref(h))
    if rc:
        # the reference constructors cannot fail; a missing device must be loud, not silent
        raise RuntimeError(_lib.last_error())
    return FlatRef(h, metric)


def flat_new_sharded(metric: str, n_shards: int, devices: Sequence[int] | None = None) -> FlatRef:
    """Additive: one index over several GPUs in this process (vb_flat_new_sharded). The handle works with
    flat_insert / flat_insert_many / flat_delete / flat_search / flat_search_batch like any other."""
    h = C.c_void_p()
    devs = (C.c_int * n_shards)(*devices) if devices is not None else None
    rc = lib().vb_flat_new_sharded(METRIC_CODE[metric], int(n_shards), devs, C.byref(h))
    if rc:
        raise RuntimeError(_lib.last_error())
    return FlatRef(h, metric)


def flat_new_l2(): return _flat_new("l2")
def flat_new_l2_squared(): return _flat_new("l2_squared")
def flat_new_cosine(): return _flat_new("cosine")
def flat_new_inner_product(): return _flat_new("inner_product")
def flat_new_negative_inner_product(): return _flat_new("negative_inner_product")
def flat_new_manhattan(): return _flat_new("manhattan")
def flat_new_chebyshev(): return _flat_new("chebyshev")
def flat_new_hamming(): return _flat_new("hamming")
def flat_new_jaccard(): return _flat_new("jaccard")


def flat_insert(index: FlatRef, id, vector):
    """nifs.rs:259-271."""
    v = _f32(vector)
    b = _enc(id)
    rc = lib().vb_flat_insert(index.handle, b, len(b), _ptr(v, _f32p), v.size)
    return _err() if rc else ("ok", ())


def flat_insert_many(index: FlatRef, vectors: Sequence[tuple]):
    """nifs.rs:273-284. ``vectors`` is a list of ``(id, vector)``."""
    ids = [v[0] for v in vectors]
    vals, off = _ragged([v[1] for v in vectors], np.float32)
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_flat_insert_many(index.handle, len(ids), blob, _ptr(ioff, _u64p), _ptr(vals, _f32p), _ptr(off, _u64p))
    return _err() if rc else ("ok", ())


def flat_insert_matrix(index: FlatRef, ids: Sequence, matrix: np.ndarray):
    """Same C entry as flat_insert_many for a dense ``[n, d]`` float32 matrix (no per-row Python objects)."""
    vals, off = _ragged(np.asarray(matrix), np.float32)
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_flat_insert_many(index.handle, len(ids), blob, _ptr(ioff, _u64p), _ptr(vals, _f32p), _ptr(off, _u64p))
    return _err() if rc else ("ok", ())


def flat_reserve(index: FlatRef, rows: int):
    """Additive: pre-size the HBM matrix for a bulk load."""
    rc = lib().vb_flat_reserve(index.handle, int(rows))
    return _err() if rc else ("ok", ())


def flat_insert_device(index: FlatRef, ids: Sequence, device_ptr: int, dimension: int):
    """Additive bulk ingest: `len(ids)` rows of `dimension` float32 at `device_ptr` (device memory, contiguous)."""
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_flat_insert_many_device(index.handle, len(ids), blob, _ptr(ioff, _u64p), C.c_void_p(int(device_ptr)),
                                          int(dimension))
    return _err() if rc else ("ok", ())


def flat_delete(index: FlatRef, id):
    """nifs.rs:286-295."""
    b = _enc(id)
    rc = lib().vb_flat_delete(index.handle, b, len(b))
    return _err() if rc else ("ok", ())


def flat_search(index: FlatRef, query, limit: int):
    """nifs.rs:297-309: ``("ok", [(id, raw)])`` ascending by (rank, id)."""
    q = _f32(query)
    h = C.c_void_p()
    rc = lib().vb_flat_search(index.handle, _ptr(q, _f32p), q.size, min(int(limit), SIZE_MAX), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def flat_search_batch(index: FlatRef, queries: np.ndarray, limit: int):
    """Additive: one call for ``[nq, d]`` queries; ``("ok", [hits per query])``."""
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
    nq, d = q.shape
    hs = (C.c_void_p * nq)()
    rc = lib().vb_flat_search_batch(index.handle, _ptr(q, _f32p), nq, d, int(limit), hs)
    return _err() if rc else ("ok", [_take_hits(C.c_void_p(h)) for h in hs])


def flat_info(index: FlatRef) -> tuple[int, int | None]:
    rows, dim = C.c_size_t(), C.c_size_t()
    lib().vb_flat_info(index.handle, C.byref(rows), C.byref(dim))
    return rows.value, (dim.value or None)


def flat_prefix_top_k(index: FlatRef, ids: Sequence | None, query, metric_code: int, dimensions: int, limit: int):
    """Additive resident form of vector_top_k (search.rs:38-73) over the index's own rows:
    ``ids=None`` scores every row, else only the listed ids (unknown ids are skipped)."""
    q = _f32(query)
    h = C.c_void_p()
    if ids is None:
        rc = lib().vb_flat_prefix_top_k(index.handle, SIZE_MAX, None, None, _ptr(q, _f32p), q.size, int(metric_code),
                                        int(dimensions), int(limit), C.byref(h))
    else:
        blob, ioff = _ids_blob(ids)
        rc = lib().vb_flat_prefix_top_k(index.handle, len(ids), blob, _ptr(ioff, _u64p), _ptr(q, _f32p), q.size,
                                        int(metric_code), int(dimensions), int(limit), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def flat_funnel_search(index: FlatRef, query, metric_code: int, stages: Sequence[int], candidates: int, limit: int):
    """Additive: the whole funnel pipeline (collection.ex:244-260) on the resident matrix."""
    q = _f32(query)
    st = (C.c_size_t * len(stages))(*[int(s) for s in stages])
    h = C.c_void_p()
    rc = lib().vb_flat_funnel_search(index.handle, _ptr(q, _f32p), q.size, int(metric_code), st, len(stages),
                                     int(candidates), int(limit), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def flat_quantized_search(index: FlatRef, query, metric_code: int, candidates: int, limit: int):
    """Additive: Hamming candidates over the resident sign codes + exact rerank (collection.ex:266-295)."""
    q = _f32(query)
    h = C.c_void_p()
    rc = lib().vb_flat_quantized_search(index.handle, _ptr(q, _f32p), q.size, int(metric_code), int(candidates),
                                        int(limit), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def vector_top_k(vectors: Sequence[tuple], query, metric_code: int, dimensions: int, limit: int):
    """nifs.rs:151-162."""
    ids = [v[0] for v in vectors]
    vals, off = _ragged([v[1] for v in vectors], np.float32)
    blob, ioff = _ids_blob(ids)
    q = _f32(query)
    h = C.c_void_p()
    rc = lib().vb_vector_top_k(len(ids), blob, _ptr(ioff, _u64p), _ptr(vals, _f32p), _ptr(off, _u64p), _ptr(q, _f32p),
                               q.size, int(metric_code), int(dimensions), min(int(limit), SIZE_MAX), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def binary_top_k(vectors: Sequence[tuple], query: Sequence[int], dimensions: int, limit: int):
    """nifs.rs:164-175."""
    ids = [v[0] for v in vectors]
    vals, off = _ragged([v[1] for v in vectors], np.uint64)
    blob, ioff = _ids_blob(ids)
    q = np.asarray([int(x) for x in query], dtype=np.uint64)
    h = C.c_void_p()
    rc = lib().vb_binary_top_k(len(ids), blob, _ptr(ioff, _u64p), _ptr(vals, _u64p), _ptr(off, _u64p), _ptr(q, _u64p),
                               q.size, int(dimensions), min(int(limit), SIZE_MAX), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def muvera_encode_batch(documents: Sequence, dimension: int, num_repetitions: int, num_simhash_projections: int, seed: int,
                        projection_dimension: int, final_projection_dimension: int | None, mode: str):
    """Additive: MUVERA fixed-dimensional encodings of MANY multi-vectors in one device
    call (muvera.rs:26-74 per document). ``mode``: "query" (sum) or "document" (average). Returns
    ``("ok", ndarray [ndocs, fde_dim])`` or ``("error", msg)``."""
    toks, doc_vec = [], np.zeros(len(documents) + 1, dtype=np.uint64)
    for i, vs in enumerate(documents):
        toks.extend(vs)
        doc_vec[i + 1] = len(toks)
    vals, off = _ragged(toks, np.float32)
    part = 1 << min(int(num_simhash_projections), 40)
    fde = int(final_projection_dimension) if final_projection_dimension else min(
        int(num_repetitions) * part * max(1, int(projection_dimension)), 16777216)
    out = np.zeros((max(1, len(documents)), max(1, fde)), dtype=np.float32)
    n = C.c_size_t()
    rc = lib().vb_muvera_encode(len(documents), _ptr(vals, _f32p), _ptr(off, _u64p), _ptr(doc_vec, _u64p), int(dimension),
                                int(num_repetitions), int(num_simhash_projections), int(seed), int(projection_dimension),
                                int(final_projection_dimension is not None), int(final_projection_dimension or 0),
                                {"query": 0, "document": 1}[mode], _ptr(out, _f32p), out.size, C.byref(n))
    if rc:
        return _err()
    return ("ok", out.reshape(-1)[: len(documents) * n.value].reshape(len(documents), n.value))


def muvera_encode_query(vectors, dimension, num_repetitions, num_simhash_projections, seed, projection_dimension,
                        final_projection_dimension):
    """nifs.rs:430-452 (one multi-vector, summed per partition): ``("ok", [float])``."""
    res = muvera_encode_batch([vectors], dimension, num_repetitions, num_simhash_projections, seed, projection_dimension,
                              final_projection_dimension, "query")
    return res if res[0] != "ok" else ("ok", res[1][0].tolist())


def muvera_encode_document(vectors, dimension, num_repetitions, num_simhash_projections, seed, projection_dimension,
                           final_projection_dimension):
    """nifs.rs:455-476 (one multi-vector, averaged per partition): ``("ok", [float])``."""
    res = muvera_encode_batch([vectors], dimension, num_repetitions, num_simhash_projections, seed, projection_dimension,
                              final_projection_dimension, "document")
    return res if res[0] != "ok" else ("ok", res[1][0].tolist())


def result_values(metric_code: int, raws, score_mode: str = "raw"):
    """Additive: ``Distance.result_values/3`` (vettore_distance.ex:525-543) over a whole hit list in one C call.
    Returns ``("ok", [(score, distance)])``."""
    r = _f32(raws)
    score, dist = np.zeros(r.size, np.float64), np.zeros(r.size, np.float64)
    _f64p = C.POINTER(C.c_double)
    rc = lib().vb_result_values(int(metric_code), {"raw": 0, "similarity": 1}[score_mode], _ptr(r, _f32p), r.size,
                                _ptr(score, _f64p), _ptr(dist, _f64p))
    return _err() if rc else ("ok", list(zip(score.tolist(), dist.tolist())))


def compress_sign_bits(vector) -> list[int]:
    """nifs.rs:125-129: bare list of u64 words."""
    v = _f32(vector)
    words = np.zeros((v.size + 63) // 64, dtype=np.uint64)
    lib().vb_compress_sign_bits(_ptr(v, _f32p), v.size, _ptr(words, _u64p))
    return [int(w) for w in words]


# ------------------------------------------------------------------ multi-vector (MaxSim)
def _tokens(docs: Sequence[Sequence[Sequence[float]]]):
    """Flattens [[token vectors] per doc] into one ragged token list + per-doc token offsets."""
    toks, doc_tok = [], np.zeros(len(docs) + 1, dtype=np.uint64)
    for i, vs in enumerate(docs):
        if isinstance(vs, np.ndarray) and vs.ndim == 2:
            toks.extend(vs)
        else:
            toks.extend(vs)
        doc_tok[i + 1] = len(toks)
    vals, off = _ragged(toks, np.float32)
    return vals, off, doc_tok


def multi_vector_score(query_vectors, document_vectors, metric_code: int):
    """nifs.rs:177-186."""
    qv, qoff = _ragged(list(query_vectors), np.float32)
    dv, doff = _ragged(list(document_vectors), np.float32)
    out = C.c_float()
    rc = lib().vb_multi_vector_score(_ptr(qv, _f32p), _ptr(qoff, _u64p), len(qoff) - 1, _ptr(dv, _f32p),
                                     _ptr(doff, _u64p), len(doff) - 1, int(metric_code), C.byref(out))
    return _err() if rc else ("ok", out.value)


def multi_vector_top_k(documents: Sequence[tuple], query_vectors, metric_code: int, limit: int):
    """nifs.rs:188-198. ``documents`` is a list of ``(id, [token vectors])``."""
    ids = [d[0] for d in documents]
    dv, doff, doc_tok = _tokens([d[1] for d in documents])
    qv, qoff = _ragged(list(query_vectors), np.float32)
    blob, ioff = _ids_blob(ids)
    h = C.c_void_p()
    rc = lib().vb_multi_vector_top_k(len(ids), blob, _ptr(ioff, _u64p), _ptr(dv, _f32p), _ptr(doff, _u64p),
                                     _ptr(doc_tok, _u64p), _ptr(qv, _f32p), _ptr(qoff, _u64p), len(qoff) - 1,
                                     int(metric_code), min(int(limit), SIZE_MAX), C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


class MvRef:
    """Handle of an HBM-resident multi-vector collection (additive; see include/vettore_b200.h)."""

    def __init__(self, metric: str, n_shards: int = 0, devices: Sequence[int] | None = None):
        h = C.c_void_p()
        if n_shards:
            devs = (C.c_int * n_shards)(*devices) if devices is not None else None
            rc = lib().vb_mv_new_sharded(METRIC_CODE[metric], int(n_shards), devs, C.byref(h))
        else:
            rc = lib().vb_mv_new(METRIC_CODE[metric], C.byref(h))
        if rc:
            raise RuntimeError(_lib.last_error())
        self._h, self.metric = h, metric

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().vb_mv_free(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h


def mv_new(metric: str) -> MvRef:
    return MvRef(metric)


def mv_new_sharded(metric: str, n_shards: int, devices: Sequence[int] | None = None) -> MvRef:
    """Additive: one multi-vector collection over several GPUs in this process (vb_mv_new_sharded)."""
    return MvRef(metric, n_shards, devices)


def mv_insert_many(index: MvRef, documents: Sequence[tuple]):
    ids = [d[0] for d in documents]
    dv, doff, doc_tok = _tokens([d[1] for d in documents])
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_mv_insert_many(index.handle, len(ids), blob, _ptr(ioff, _u64p), _ptr(dv, _f32p), _ptr(doff, _u64p),
                                 _ptr(doc_tok, _u64p))
    return _err() if rc else ("ok", ())


def mv_insert_tensor(index: MvRef, ids: Sequence, tokens: np.ndarray):
    """Dense ``[ndocs, tokens_per_doc, dim]`` float32 ingest through the same C entry."""
    t = np.ascontiguousarray(tokens, dtype=np.float32)
    nd, tp, d = t.shape
    vals = t.reshape(-1)
    off = np.arange(nd * tp + 1, dtype=np.uint64) * np.uint64(d)
    doc_tok = np.arange(nd + 1, dtype=np.uint64) * np.uint64(tp)
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_mv_insert_many(index.handle, nd, blob, _ptr(ioff, _u64p), _ptr(vals, _f32p), _ptr(off, _u64p),
                                 _ptr(doc_tok, _u64p))
    return _err() if rc else ("ok", ())


def mv_reserve(index: MvRef, docs: int, tokens: int, dimension: int):
    rc = lib().vb_mv_reserve(index.handle, int(docs), int(tokens), int(dimension))
    return _err() if rc else ("ok", ())


def mv_insert_device(index: MvRef, ids: Sequence, device_ptr: int, tokens_per_doc: int, dimension: int):
    """Uniform documents whose tokens are already in device memory (``device_ptr`` = address of a
    row-major ``[len(ids) * tokens_per_doc, dimension]`` float32 matrix)."""
    blob, ioff = _ids_blob(ids)
    rc = lib().vb_mv_insert_many_device(index.handle, len(ids), blob, _ptr(ioff, _u64p), C.c_void_p(device_ptr),
                                        int(tokens_per_doc), int(dimension))
    return _err() if rc else ("ok", ())


def mv_insert_ragged_device(index: MvRef, ids: Sequence, device_ptr: int, doc_tok, dimension: int):
    """Ragged documents whose tokens are already in device memory: document ``i`` owns rows
    ``[doc_tok[i], doc_tok[i + 1])`` of the row-major ``[tokens, dimension]`` float32 matrix at ``device_ptr``."""
    blob, ioff = _ids_blob(ids)
    dt = np.ascontiguousarray(doc_tok, dtype=np.uint64)
    rc = lib().vb_mv_insert_ragged_device(index.handle, len(ids), blob, _ptr(ioff, _u64p), C.c_void_p(device_ptr),
                                          _ptr(dt, _u64p), int(dimension))
    return _err() if rc else ("ok", ())


def mv_delete(index: MvRef, id):
    b = _enc(id)
    rc = lib().vb_mv_delete(index.handle, b, len(b))
    return _err() if rc else ("ok", ())


def mv_search(index: MvRef, query_vectors, limit: int):
    qv, qoff = _ragged(list(query_vectors) if not isinstance(query_vectors, np.ndarray) else query_vectors, np.float32)
    h = C.c_void_p()
    rc = lib().vb_mv_search(index.handle, _ptr(qv, _f32p), _ptr(qoff, _u64p), len(qoff) - 1, min(int(limit), SIZE_MAX),
                            C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def mv_search_packed_device(index: MvRef, query_vectors, limit: int, d_keys, d_values, d_rows, d_counts):
    """Additive (document-sharded MaxSim): `mv_search` whose sorted top-k also stays on the device."""
    qv, qoff = _ragged(list(query_vectors) if not isinstance(query_vectors, np.ndarray) else query_vectors, np.float32)
    h = C.c_void_p()
    rc = lib().vb_mv_search_packed_device(index.handle, _ptr(qv, _f32p), _ptr(qoff, _u64p), len(qoff) - 1,
                                          min(int(limit), SIZE_MAX), d_keys, d_values, d_rows, d_counts, C.byref(h))
    return _err() if rc else ("ok", _take_hits(h))


def mv_info(index: MvRef):
    docs, toks, dim = C.c_size_t(), C.c_size_t(), C.c_size_t()
    lib().vb_mv_info(index.handle, C.byref(docs), C.byref(toks), C.byref(dim))
    return docs.value, toks.value, (dim.value or None)
