"""ctypes loader for libvettore_b200.so (the C ABI of include/vettore_b200.h).

Fails loudly: a missing library or a missing CUDA device is an error, never a fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libvettore_b200.so")

VB_OK, VB_ERR, VB_ERR_CUDA = 0, 1, 2
SIZE_MAX = C.c_size_t(-1).value

_sz, _u64p, _u32p, _f32p = C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
_vp, _vpp = C.c_void_p, C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/vettore_b200.h declaration by declaration.
SIGNATURES = {
    "vb_last_error": (C.c_char_p, []),
    "vb_version": (C.c_char_p, []),
    "vb_device_count": (C.c_int, []),
    "vb_hits_len": (_sz, [_vp]),
    "vb_hits_id": (C.POINTER(C.c_char), [_vp, _sz, C.POINTER(_sz)]),
    "vb_hits_value": (C.c_float, [_vp, _sz]),
    "vb_hits_index": (C.c_uint64, [_vp, _sz]),
    "vb_hits_export": (_sz, [_vp, C.POINTER(C.c_void_p), C.POINTER(_u64p), C.POINTER(_f32p), C.POINTER(_u64p)]),
    "vb_hits_free": (None, [_vp]),
    "vb_flat_new": (C.c_int, [C.c_int, _vpp]),
    "vb_flat_new_sharded": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), _vpp]),
    "vb_flat_free": (None, [_vp]),
    "vb_flat_insert": (C.c_int, [_vp, C.c_char_p, _sz, _f32p, _sz]),
    "vb_flat_insert_many": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _f32p, _u64p]),
    "vb_flat_reserve": (C.c_int, [_vp, _sz]),
    "vb_flat_insert_many_device": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _vp, _sz]),
    "vb_flat_delete": (C.c_int, [_vp, C.c_char_p, _sz]),
    "vb_flat_search": (C.c_int, [_vp, _f32p, _sz, _sz, _vpp]),
    "vb_flat_search_batch": (C.c_int, [_vp, _f32p, _sz, _sz, _sz, _vpp]),
    "vb_flat_info": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "vb_flat_prefix_top_k": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _f32p, _sz, C.c_int, _sz, _sz, _vpp]),
    "vb_flat_funnel_search": (C.c_int, [_vp, _f32p, _sz, C.c_int, C.POINTER(_sz), _sz, _sz, _sz, _vpp]),
    "vb_flat_quantized_search": (C.c_int, [_vp, _f32p, _sz, C.c_int, _sz, _sz, _vpp]),
    "vb_flat_search_device": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_flat_hamming_device": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_flat_rerank_owned_device": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _vp, _sz, C.c_uint32, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_flat_device_status": (C.c_int, [_vp, _u32p]),
    "vb_flat_set_id_ranks": (C.c_int, [_vp, _u32p, _sz]),
    "vb_topk_merge_device": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_peer_new": (C.c_int, [C.c_int, C.c_int, _sz, _vpp, _vp]),
    "vb_peer_free": (None, [_vp]),
    "vb_peer_connect_ipc": (C.c_int, [_vp, _vp]),
    "vb_peer_connect_local": (C.c_int, [_vpp, C.c_int]),
    "vb_peer_exchange_merge": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_peer_push": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz, _vp]),
    "vb_peer_wait_merge": (C.c_int, [_vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "vb_peer_error": (C.c_int, [_vp, _u32p]),
    "vb_vector_top_k": (C.c_int, [_sz, C.c_char_p, _u64p, _f32p, _u64p, _f32p, _sz, C.c_int, _sz, _sz, _vpp]),
    "vb_binary_top_k": (C.c_int, [_sz, C.c_char_p, _u64p, _u64p, _u64p, _u64p, _sz, _sz, _sz, _vpp]),
    "vb_muvera_encode": (C.c_int, [_sz, _f32p, _u64p, _u64p, _sz, _sz, _sz, C.c_uint64, _sz, C.c_int, _sz, C.c_int, _f32p, _sz,
                                   C.POINTER(_sz)]),
    "vb_result_values": (C.c_int, [C.c_int, C.c_int, _f32p, _sz, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vb_compress_sign_bits": (C.c_int, [_f32p, _sz, _u64p]),
    "vb_multi_vector_top_k": (C.c_int, [_sz, C.c_char_p, _u64p, _f32p, _u64p, _u64p, _f32p, _u64p, _sz, C.c_int, _sz, _vpp]),
    "vb_multi_vector_score": (C.c_int, [_f32p, _u64p, _sz, _f32p, _u64p, _sz, C.c_int, _f32p]),
    "vb_mv_new": (C.c_int, [C.c_int, _vpp]),
    "vb_mv_new_sharded": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), _vpp]),
    "vb_mv_free": (None, [_vp]),
    "vb_mv_insert_many": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _f32p, _u64p, _u64p]),
    "vb_mv_reserve": (C.c_int, [_vp, _sz, _sz, _sz]),
    "vb_mv_insert_many_device": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _vp, _sz, _sz]),
    "vb_mv_insert_ragged_device": (C.c_int, [_vp, _sz, C.c_char_p, _u64p, _vp, _u64p, _sz]),
    "vb_mv_delete": (C.c_int, [_vp, C.c_char_p, _sz]),
    "vb_mv_search": (C.c_int, [_vp, _f32p, _u64p, _sz, _sz, _vpp]),
    "vb_mv_search_packed_device": (C.c_int, [_vp, _f32p, _u64p, _sz, _sz, _vp, _vp, _vp, _vp, _vpp]),
    "vb_mv_set_id_ranks": (C.c_int, [_vp, _u32p, _sz]),
    "vb_mv_info": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads the shared library (once). Raises LibraryMissing when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LibraryMissing(
                f"{SO_PATH} not found: build it with `python -m vettore_b200.build` "
                "(vettore_b200 has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().vb_last_error().decode("utf-8", "replace")
