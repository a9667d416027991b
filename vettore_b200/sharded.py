"""Row-sharded flat search over the GPUs of one box: one process per GPU
(``torch.distributed``), each rank owns a contiguous row range of the corpus in its own
HBM-resident ``vb_flat`` index; a query is replicated, every rank runs the fused
scan + top-k kernel on its shard, the per-rank sorted top-k records are exchanged with ONE
all-gather over NVLink (NCCL; gloo on CPU for the host-logic tests), and every rank runs
the K7 merge kernel to select the global top-k. SURVEY.md §8(e).

Cross-shard tie-breaks need globally comparable id ranks: ``set_global_ranks`` installs
them (for ids that sort like the global row number the rank is just base + row).

PyTorch is plumbing here: device buffers, streams, the process group. The scan and the
merge are the library's own kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, nifs
from ._lib import lib


def packed_layout(nq: int, k: int) -> dict:
    """Byte offsets of one shard's record: keys[nq][k] u64 | values f32 | rows u32 | counts[nq] u32."""
    keys = 0
    values = keys + nq * k * 8
    rows = values + nq * k * 4
    counts = rows + nq * k * 4
    total = counts + nq * 4
    total = (total + 15) // 16 * 16
    return {"keys": keys, "values": values, "rows": rows, "counts": counts, "bytes": total}


def merged_offsets(nq: int, k_out: int) -> tuple[int, int, int, int]:
    """Byte offsets (keys u64, values f32, rows u64 = shard << 32 | row, counts u32) of a merged record."""
    o_keys = 0
    o_vals = o_keys + nq * k_out * 8
    o_rows = (o_vals + nq * k_out * 4 + 7) // 8 * 8
    o_counts = o_rows + nq * k_out * 8
    return o_keys, o_vals, o_rows, o_counts


class Exchange:
    """Exchange + final select of one packed top-k record per rank.

    ``peer`` (default on CUDA): the library's own kernel stores the record into every peer's buffer over
    NVLink peer memory (buffers mapped across the processes with CUDA IPC handles, exchanged once over the
    process group), publishes a flag, waits for the peers' flags and runs the K7 select — one launch per
    step, no collective call. ``nccl`` (``VB_EXCHANGE=nccl``, or when the peer mapping cannot be set up):
    one ``all_gather_into_tensor`` + the K7 merge kernel. Both produce the same merged record."""

    def __init__(self, lay: dict, nq: int, k_in: int, k_out: int, group, device: torch.device):
        import os

        self.lay, self.nq, self.k_in, self.k_out, self.group, self.device = lay, nq, k_in, k_out, group, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.offsets = merged_offsets(nq, k_out)
        self.peer = None
        self.gathered = None
        self.mode = "none"
        if self.world == 1:
            return
        self.mode = "nccl"
        if device.type == "cuda" and os.environ.get("VB_EXCHANGE", "peer") != "nccl" and self.world <= 8:
            self._try_peer()
        if self.mode == "nccl":
            self.gathered = torch.zeros(self.world * lay["bytes"], dtype=torch.uint8, device=device)

    def _try_peer(self):
        handle = (C.c_ubyte * 64)()
        px = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib().vb_peer_new(self.world, self.rank, self.lay["bytes"], C.byref(px), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if rc == 0 else None, group=self.group)
        ok = rc == 0 and all(h is not None for h in handles)
        if ok:
            blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
            with torch.cuda.device(self.device):
                ok = lib().vb_peer_connect_ipc(px, blob) == 0
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.peer, self.mode = px, "peer"
        elif rc == 0:
            lib().vb_peer_free(px)

    @property
    def name(self) -> str:
        return {"none": "none (single shard)", "nccl": "NCCL all-gather + K7 merge kernel",
                "peer": "NVLink peer-memory stores + fused wait/select kernel (vb_peer_exchange_merge)"}[self.mode]

    @property
    def launches(self) -> int:
        return {"none": 0, "nccl": 2, "peer": 1 if self.nq <= 4 else 2}[self.mode]

    def run(self, local: torch.Tensor, out: torch.Tensor, stream: C.c_void_p) -> tuple[int, int, int, int]:
        """``local``: this rank's packed record (device). Writes the merged record into ``out``; returns its offsets."""
        lay = self.lay
        o_keys, o_vals, o_rows, o_counts = self.offsets
        p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
        if self.mode == "peer":
            rc = lib().vb_peer_exchange_merge(self.peer, p(local), self.nq, self.k_in, self.k_out, lay["keys"], lay["values"],
                                              lay["rows"], lay["counts"], p(out, o_keys), p(out, o_vals), p(out, o_rows),
                                              p(out, o_counts), stream)
            if rc:
                raise RuntimeError(_lib.last_error())
            return self.offsets
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, local, group=self.group)
            src = self.gathered
        else:
            src = local
        return _merge_gathered(src, lay, self.nq, self.world, self.k_in, self.k_out, out, stream)

    def __del__(self):
        px, self.peer = getattr(self, "peer", None), None
        if px:
            try:
                lib().vb_peer_free(px)
            except Exception:
                pass


@dataclass
class ShardHit:
    shard: int
    row: int
    value: float


class ShardedFlat:
    """One rank's view of a row-sharded flat index."""

    def __init__(self, index: nifs.FlatRef, k: int, nq: int = 1, group=None, device: torch.device | None = None):
        self.index = index
        self.k, self.nq = int(k), int(nq)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.layout = packed_layout(self.nq, self.k)
        nbytes = self.layout["bytes"]
        self.local = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(nbytes + self.nq * self.k * 4 + 16, dtype=torch.uint8, device=self.device)
        self.exchange = Exchange(self.layout, self.nq, self.k, self.k, group, self.device)
        self.exchange_name = self.exchange.name
        self.launches_per_search = 2 + self.exchange.launches  # scan, unpack (+ exchange / merge when sharded)

    # ---- device-side pieces ------------------------------------------------------------
    def _ptr(self, t: torch.Tensor, off: int = 0) -> C.c_void_p:
        return C.c_void_p(t.data_ptr() + off)

    def search_device(self, d_queries: torch.Tensor) -> torch.Tensor:
        """Queries ``[nq, q_stride]`` float32 on the device -> packed global top-k record
        (layout of ``packed_layout`` but rows are u64 ``shard << 32 | row``) on the device.
        Enqueued on the current torch stream; no host synchronisation."""
        lay = self.layout
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nq, stride = d_queries.shape
        rc = lib().vb_flat_search_device(self.index.handle, self._ptr(d_queries), nq, stride, self.k,
                                         self._ptr(self.local, lay["keys"]), self._ptr(self.local, lay["values"]),
                                         self._ptr(self.local, lay["rows"]), self._ptr(self.local, lay["counts"]),
                                         stream)
        if rc:
            raise RuntimeError(_lib.last_error())
        if self.world == 1:
            return self.local  # single shard: the local record already is the global top-k
        o_keys, o_vals, o_rows, o_counts = self.exchange.run(self.local, self.out, stream)
        self._out_off = (o_keys, o_vals, o_rows, o_counts)
        return self.out

    def decode(self, out_host: np.ndarray, nq: int | None = None) -> list[list[ShardHit]]:
        nq = nq or self.nq
        if self.world == 1:
            lay = self.layout
            vals = out_host[lay["values"]:lay["values"] + nq * self.k * 4].view(np.float32).reshape(nq, self.k)
            rows = out_host[lay["rows"]:lay["rows"] + nq * self.k * 4].view(np.uint32).reshape(nq, self.k)
            counts = out_host[lay["counts"]:lay["counts"] + nq * 4].view(np.uint32)
            return [[ShardHit(0, int(rows[q, i]), float(vals[q, i])) for i in range(int(counts[q]))]
                    for q in range(nq)]
        _, o_vals, o_rows, o_counts = self._out_off
        vals = out_host[o_vals:o_vals + nq * self.k * 4].view(np.float32).reshape(nq, self.k)
        rows = out_host[o_rows:o_rows + nq * self.k * 8].view(np.uint64).reshape(nq, self.k)
        counts = out_host[o_counts:o_counts + nq * 4].view(np.uint32)
        res = []
        for q in range(nq):
            res.append([ShardHit(int(rows[q, i]) >> 32, int(rows[q, i]) & 0xFFFFFFFF, float(vals[q, i]))
                        for i in range(int(counts[q]))])
        return res

    # ---- host-facing search (the e2e path: host query in, host hits out) ----------------
    def search(self, queries_host: torch.Tensor) -> list[list[ShardHit]]:
        """``queries_host``: pinned ``[nq, q_stride]`` float32. Includes the H2D copy of the
        query and the D2H read of the result."""
        dq = queries_host.to(self.device, non_blocking=True)
        out = self.search_device(dq)
        host = out.cpu()
        check_device_status(self.index)
        return self.decode(host.numpy(), dq.shape[0])


def check_device_status(index: nifs.FlatRef) -> None:
    """The stream-ordered device entries cannot return the reference's "metric overflow" (flat.rs:105): they
    raise a sticky bit instead, surfaced here once the result is on the host."""
    st = C.c_uint32(0)
    rc = lib().vb_flat_device_status(index.handle, C.byref(st))
    if rc:
        raise RuntimeError(_lib.last_error())
    if st.value & 1:
        raise RuntimeError("metric overflow")


def _merge_gathered(gathered: torch.Tensor, lay: dict, nq: int, lists: int, k_in: int, k_out: int, out: torch.Tensor,
                    stream: C.c_void_p) -> tuple[int, int, int, int]:
    """K7 over `lists` all-gathered records of layout `lay`; returns the byte offsets (keys, values, rows u64,
    counts) of the merged record inside `out`."""
    o_keys = 0
    o_vals = o_keys + nq * k_out * 8
    o_rows = (o_vals + nq * k_out * 4 + 7) // 8 * 8
    o_counts = o_rows + nq * k_out * 8
    base = gathered.data_ptr()
    rc = lib().vb_topk_merge_device(C.c_void_p(base + lay["keys"]), C.c_void_p(base + lay["values"]),
                                    C.c_void_p(base + lay["rows"]), C.c_void_p(base + lay["counts"]), lay["bytes"],
                                    nq, lists, k_in, k_out, C.c_void_p(out.data_ptr() + o_keys),
                                    C.c_void_p(out.data_ptr() + o_vals), C.c_void_p(out.data_ptr() + o_rows),
                                    C.c_void_p(out.data_ptr() + o_counts), stream)
    if rc:
        raise RuntimeError(_lib.last_error())
    return o_keys, o_vals, o_rows, o_counts


def _decode_merged(host: np.ndarray, offs: tuple[int, int, int, int], k: int) -> list[ShardHit]:
    _, o_vals, o_rows, o_counts = offs
    vals = host[o_vals:o_vals + k * 4].view(np.float32)
    rows = host[o_rows:o_rows + k * 8].view(np.uint64)
    count = int(host[o_counts:o_counts + 4].view(np.uint32)[0])
    return [ShardHit(int(rows[i]) >> 32, int(rows[i]) & 0xFFFFFFFF, float(vals[i])) for i in range(count)]


class ShardedQuantized:
    """Row-sharded ``quantized_search`` (collection.ex:699-713): every rank scans
    the sign codes of ITS rows for its best ``candidates`` (K3), the lists are all-gathered and merged into
    the global candidate set (K7), every rank reranks exactly the survivors it owns (K4, vector_top_k
    semantics), and a second, tiny all-gather + merge yields the global top-``limit``."""

    def __init__(self, index: nifs.FlatRef, candidates: int, limit: int, metric_code: int, group=None,
                 device: torch.device | None = None):
        self.index, self.cand, self.k, self.metric_code = index, int(candidates), int(limit), int(metric_code)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.lay_c, self.lay_k = packed_layout(1, self.cand), packed_layout(1, self.k)
        z = lambda n: torch.zeros(n, dtype=torch.uint8, device=self.device)
        self.local_c, self.local_k = z(self.lay_c["bytes"]), z(self.lay_k["bytes"])
        self.out_c, self.out_k = z(self.cand * 20 + 64), z(self.k * 20 + 64)
        self.ex_c = Exchange(self.lay_c, 1, self.cand, self.cand, group, self.device)
        self.ex_k = Exchange(self.lay_k, 1, self.k, self.k, group, self.device)
        self.exchange_name = self.ex_c.name

    def candidates_device(self, d_query: torch.Tensor) -> tuple[torch.Tensor, tuple[int, int, int, int]]:
        """Stage 1 only: this shard's Hamming candidates (K6 + K3), the exchange and the global select."""
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
        _, stride = d_query.shape
        lc = self.lay_c
        rc = lib().vb_flat_hamming_device(self.index.handle, p(d_query), 1, stride, self.cand, p(self.local_c, lc["keys"]),
                                          p(self.local_c, lc["values"]), p(self.local_c, lc["rows"]),
                                          p(self.local_c, lc["counts"]), stream)
        if rc:
            raise RuntimeError(_lib.last_error())
        return self.out_c, self.ex_c.run(self.local_c, self.out_c, stream)

    def search_device(self, d_query: torch.Tensor) -> tuple[torch.Tensor, tuple[int, int, int, int]]:
        """``d_query``: ``[1, q_stride]`` float32 on the device. Returns the merged record and its offsets."""
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
        _, stride = d_query.shape
        lk = self.lay_k
        _, offs_c = self.candidates_device(d_query)
        rc = lib().vb_flat_rerank_owned_device(self.index.handle, p(d_query), stride, self.metric_code,
                                               p(self.out_c, offs_c[2]), p(self.out_c, offs_c[3]), self.cand, self.rank,
                                               self.k, p(self.local_k, lk["keys"]), p(self.local_k, lk["values"]),
                                               p(self.local_k, lk["rows"]), p(self.local_k, lk["counts"]), stream)
        if rc:
            raise RuntimeError(_lib.last_error())
        offs_k = self.ex_k.run(self.local_k, self.out_k, stream)
        return self.out_k, offs_k

    def search(self, query_host: torch.Tensor) -> list[ShardHit]:
        out, offs = self.search_device(query_host.to(self.device, non_blocking=True))
        host = out.cpu().numpy()
        check_device_status(self.index)
        return _decode_merged(host, offs, self.k)


class ShardedMv:
    """Document-sharded MaxSim (multi_vector.rs:90-132): every rank scores the
    documents it owns (K5) and keeps its sorted top-k on the device; one all-gather + K7 merge."""

    def __init__(self, index: "nifs.MvRef", k: int, group=None, device: torch.device | None = None):
        self.index, self.k, self.group = index, int(k), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.lay = packed_layout(1, self.k)
        self.local = torch.zeros(self.lay["bytes"], dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(self.k * 20 + 64, dtype=torch.uint8, device=self.device)
        self.exchange = Exchange(self.lay, 1, self.k, self.k, group, self.device)
        self.exchange_name = self.exchange.name
        self._t = [0.0, 0.0, 0]   # seconds in the local scan call / in exchange + read-back, calls

    def phase_ms(self) -> dict:
        """Mean host-clock milliseconds per search spent in the local scan call and in exchange + D2H."""
        n = max(1, self._t[2])
        return {"local_scan_call": 1e3 * self._t[0] / n, "exchange_and_readback": 1e3 * self._t[1] / n, "calls": self._t[2]}

    def search(self, query_tokens: np.ndarray) -> list[ShardHit]:
        """``query_tokens``: host ``[tq, dim]`` float32 (the reference API takes the query by value)."""
        import time as _time
        lay = self.lay
        p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
        t0 = _time.perf_counter()
        res = nifs.mv_search_packed_device(self.index, query_tokens, self.k, p(self.local, lay["keys"]),
                                           p(self.local, lay["values"]), p(self.local, lay["rows"]),
                                           p(self.local, lay["counts"]))
        if res[0] != "ok":
            raise RuntimeError(res[1])
        torch.cuda.synchronize(self.device)   # the library scored on its own stream
        t1 = _time.perf_counter()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        offs = self.exchange.run(self.local, self.out, stream)
        host = self.out.cpu().numpy()
        t2 = _time.perf_counter()
        self._t[0] += t1 - t0
        self._t[1] += t2 - t1
        self._t[2] += 1
        return _decode_merged(host, offs, self.k)


def set_global_mv_ranks(index: "nifs.MvRef", base: int, docs: int) -> None:
    """Id ranks for document ids that sort like the global document number."""
    ranks = (np.arange(docs, dtype=np.uint64) + np.uint64(base)).astype(np.uint32)
    rc = lib().vb_mv_set_id_ranks(index.handle, ranks.ctypes.data_as(C.POINTER(C.c_uint32)), docs)
    if rc:
        raise RuntimeError(_lib.last_error())


def set_global_ranks(index: nifs.FlatRef, base: int, rows: int) -> None:
    """Id ranks for ids that sort like the global row number (zero-padded decimals)."""
    ranks = (np.arange(rows, dtype=np.uint64) + np.uint64(base)).astype(np.uint32)
    rc = lib().vb_flat_set_id_ranks(index.handle, ranks.ctypes.data_as(C.POINTER(C.c_uint32)), rows)
    if rc:
        raise RuntimeError(_lib.last_error())
