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
        self.gathered = torch.zeros(self.world * nbytes, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(nbytes + self.nq * self.k * 4 + 16, dtype=torch.uint8, device=self.device)
        self.launches_per_search = 2 + (1 if self.world > 1 else 0)  # scan, unpack (+ merge when sharded)

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
        dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        src, lists = self.gathered, self.world
        o_keys = 0
        o_vals = o_keys + nq * self.k * 8
        o_rows = o_vals + nq * self.k * 4
        o_rows = (o_rows + 7) // 8 * 8
        o_counts = o_rows + nq * self.k * 8
        rc = lib().vb_topk_merge_device(self._ptr(src, lay["keys"]), self._ptr(src, lay["values"]),
                                        self._ptr(src, lay["rows"]), self._ptr(src, lay["counts"]), lay["bytes"],
                                        nq, lists, self.k, self.k, self._ptr(self.out, o_keys),
                                        self._ptr(self.out, o_vals), self._ptr(self.out, o_rows),
                                        self._ptr(self.out, o_counts), stream)
        if rc:
            raise RuntimeError(_lib.last_error())
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
        return self.decode(host.numpy(), dq.shape[0])


def set_global_ranks(index: nifs.FlatRef, base: int, rows: int) -> None:
    """Id ranks for ids that sort like the global row number (zero-padded decimals)."""
    ranks = (np.arange(rows, dtype=np.uint64) + np.uint64(base)).astype(np.uint32)
    rc = lib().vb_flat_set_id_ranks(index.handle, ranks.ctypes.data_as(C.POINTER(C.c_uint32)), rows)
    if rc:
        raise RuntimeError(_lib.last_error())
