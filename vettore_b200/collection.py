"""Scan-path pipelines of ``Vettore.Collection`` (reference lib/vettore/collection.ex) on
top of the HBM-resident indexes — the re-routing SURVEY.md §8(f) rank 2 describes:
``search`` / ``funnel_search`` / ``quantized_search`` / ``multi_vector_search`` keep their
option names, defaults and result shaping, but instead of ``store.all`` + by-value NIFs per
call (collection.ex:254, 284, 320) they run on the device-resident mirrors.

The canonical record store stays on the host (a dict here, ETS in the reference): the
device holds only ids + vectors, exactly like ``Vettore.Index.Flat`` (index/flat.ex:1-8);
hits whose id is gone from the store are dropped (index/flat.ex:72-91).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

from . import nifs

MAX_NIF_USIZE = 4_294_967_295  # index/flat.ex:13
SIMILARITY_METRICS = ("cosine", "inner_product")


@dataclass
class Embedding:  # lib/vettore_embedding.ex
    id: str
    vector: Sequence[float]
    value: Any = None
    vectors: Sequence[Sequence[float]] | None = None
    binary_vector: Sequence[int] | None = None
    metadata: dict = field(default_factory=dict)


@dataclass
class Result:  # lib/vettore/result.ex
    id: str
    value: Any
    score: float
    distance: float | None
    metric: str
    metadata: dict


def normalize_l2(v) -> np.ndarray:
    """distances.rs:350-361: divide by the f64 norm, cast to f32; zero stays zero."""
    a = np.asarray(v, dtype=np.float32)
    n = np.sqrt(np.sum(a.astype(np.float64) ** 2))
    return a.copy() if n == 0.0 else (a.astype(np.float64) / n).astype(np.float32)


def result_values(metric: str, raw: float, score_mode: str = "raw"):
    """vettore_distance.ex:525-548."""
    if metric == "negative_inner_product":
        return -raw, raw
    if metric in SIMILARITY_METRICS:
        dist = 1.0 - raw if metric == "cosine" else -raw
        if score_mode == "raw":
            return raw, dist
        return ((raw + 1.0) / 2.0 if metric == "cosine" else raw), dist
    if score_mode == "raw":
        return -raw, raw
    return 1.0 / (1.0 + raw), raw


class Collection:
    def __init__(self, metric: str = "cosine", dimensions: int | None = None, normalize: str | None = None,
                 score: str = "raw"):
        if metric not in nifs.METRIC_CODE:
            raise ValueError(f"unsupported metric {metric}")
        self.metric, self.dimensions, self.score = metric, dimensions, score
        self.normalize = normalize if normalize is not None else ("l2" if metric == "cosine" else "none")
        self.store: dict[str, Embedding] = {}
        self.index = getattr(nifs, f"flat_new_{metric}")()
        self._mv: nifs.MvRef | None = None
        self._mv_dirty = True

    # ---- ingest ---------------------------------------------------------------------------
    def _prepare(self, vector) -> np.ndarray:
        v = np.asarray(vector, dtype=np.float32)
        return normalize_l2(v) if self.normalize == "l2" else v

    def put_many(self, embeddings: Sequence[Embedding]):
        prepared = []
        for e in embeddings:
            v = self._prepare(e.vector)
            if self.dimensions is None:
                self.dimensions = int(v.size)
            vs = [self._prepare(t) for t in e.vectors] if e.vectors else None
            prepared.append(Embedding(e.id, v, e.value, vs, e.binary_vector, e.metadata))
        res = nifs.flat_insert_many(self.index, [(e.id, e.vector) for e in prepared])   # index/flat.ex:35-39
        if res[0] != "ok":
            return res
        for e in prepared:
            self.store[e.id] = e
        self._mv_dirty = True
        return ("ok", ())

    def put(self, embedding: Embedding):
        return self.put_many([embedding])

    def delete(self, id: str):
        self.store.pop(id, None)
        self._mv_dirty = True
        return nifs.flat_delete(self.index, id)

    # ---- helpers --------------------------------------------------------------------------
    @staticmethod
    def _validate_limit(limit):
        return isinstance(limit, int) and 0 < limit <= MAX_NIF_USIZE

    def _results(self, hits):
        """index/flat.ex:72-91 for the whole hit list: ONE vb_result_values call shapes every raw value
        (SURVEY.md §8(f) rank 3), then the store lookups; ids the store no longer has are dropped."""
        st, shaped = nifs.result_values(nifs.METRIC_CODE[self.metric], [raw for _, raw in hits], self.score)
        assert st == "ok", shaped
        out = []
        for (id_, _raw), (score, dist) in zip(hits, shaped):
            e = self.store.get(id_)
            if e is None:
                continue
            out.append(Result(id_, e.value, score, dist, self.metric, e.metadata))
        return out

    def _candidates(self, candidates, limit):
        return max(limit * 10, limit) if candidates is None else candidates   # collection.ex:509-510

    # ---- searches -------------------------------------------------------------------------
    def search(self, query, limit: int = 10):
        """Vettore.search/3 with index: :flat (collection.ex:224-228 -> index/flat.ex:49-57)."""
        if not self._validate_limit(limit):
            return ("error", "invalid_limit")
        st, hits = nifs.flat_search(self.index, self._prepare(query), limit)
        return (st, hits) if st != "ok" else ("ok", self._results(hits))

    def funnel_search(self, query, limit: int = 10, candidates: int | None = None, stages: Sequence[int] | None = None,
                      dimensions: int | None = None):
        """collection.ex:244-260."""
        if not self._validate_limit(limit):
            return ("error", "invalid_limit")
        candidates = self._candidates(candidates, limit)
        if not (self._validate_limit(candidates) and candidates >= limit):
            return ("error", "invalid_candidates")
        if stages is None:
            stages = [dimensions] if dimensions is not None else [min(self.dimensions or 1, 128)]   # :660-672
        code = nifs.METRIC_CODE[self.metric]
        st, hits = nifs.flat_funnel_search(self.index, self._prepare(query), code, list(stages), candidates, limit)
        return (st, hits) if st != "ok" else ("ok", self._results(hits))

    def quantized_search(self, query, limit: int = 10, candidates: int | None = None):
        """collection.ex:266-295."""
        if not self._validate_limit(limit):
            return ("error", "invalid_limit")
        candidates = self._candidates(candidates, limit)
        if not (self._validate_limit(candidates) and candidates >= limit):
            return ("error", "invalid_candidates")
        code = nifs.METRIC_CODE[self.metric]
        st, hits = nifs.flat_quantized_search(self.index, self._prepare(query), code, candidates, limit)
        return (st, hits) if st != "ok" else ("ok", self._results(hits))

    def _ensure_mv(self):
        if self._mv is None or self._mv_dirty:
            self._mv = nifs.mv_new(self.metric)
            docs = [(e.id, e.vectors if e.vectors else [e.vector]) for e in self.store.values()]   # collection.ex:773-777
            if docs:
                res = nifs.mv_insert_many(self._mv, docs)
                if res[0] != "ok":
                    return res
            self._mv_dirty = False
        return ("ok", ())

    def multi_vector_search(self, query_vectors, limit: int = 10):
        """collection.ex:313-323; results carry score only (collection.ex:807-817)."""
        if not self._validate_limit(limit):
            return ("error", "invalid_limit")
        res = self._ensure_mv()
        if res[0] != "ok":
            return res
        qs = [self._prepare(q) for q in query_vectors]
        st, hits = nifs.mv_search(self._mv, qs, limit)
        if st != "ok":
            return (st, hits)
        out = []
        for id_, s in hits:
            e = self.store.get(id_)
            if e is not None:
                out.append(Result(id_, e.value, s, None, self.metric, e.metadata))
        return ("ok", out)
