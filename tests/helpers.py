"""Shared helpers for parity tests: the reference's own comparison rules."""
import struct

import numpy as np

METRICS = ["l2", "l2_squared", "cosine", "inner_product", "negative_inner_product",
           "manhattan", "chebyshev", "hamming", "jaccard"]

TOL = 1e-5  # BASELINE.json north_star: float scores within 1e-5 relative


def total_order_key(x: float) -> int:
    """f32::total_cmp as an ascending unsigned key."""
    b = struct.unpack("<I", struct.pack("<f", np.float32(x)))[0]
    return (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b ^ 0x80000000)


def close(a: float, b: float, tol: float = TOL) -> bool:
    """distances.rs:487-493 assert_close: |a-b| <= tol * max(1, |a|, |b|)."""
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


def assert_hits_match(actual, expected, tol: float = TOL, exact_ids: bool = False):
    """Hits are [(id, value)]. Values within `tol`; id order equal except for ties inside
    the tolerance (north_star's parity bar). With exact_ids the id lists must be equal."""
    assert len(actual) == len(expected), (len(actual), len(expected))
    for (ia, va), (ie, ve) in zip(actual, expected):
        assert close(va, ve, tol), (ia, va, ie, ve)
    ids_a, ids_e = [h[0] for h in actual], [h[0] for h in expected]
    if exact_ids:
        assert ids_a == ids_e
        return
    if ids_a == ids_e:
        return
    # Allowed difference: permutations inside runs of values that tie within tolerance, and
    # swaps across the cut-off boundary with a value tying the last expected value.
    exp_val = dict(expected)
    last = expected[-1][1] if expected else 0.0
    for pos, (ia, va) in enumerate(actual):
        ie, ve = expected[pos]
        if ia == ie:
            continue
        assert close(va, ve, tol), f"order differs outside tolerance at {pos}: {ia}={va} vs {ie}={ve}"
        if ia not in exp_val:
            assert close(va, last, tol), f"{ia}={va} not in expected and not a boundary tie ({last})"
