#!/usr/bin/env python
"""Secondary benchmark: K3 sign-code Hamming scan (BASELINE.json configs[3] shape: N x 1024-bit codes,
1000 candidates). Codes are generated on the device. One JSON line."""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vettore_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--dims", type=int, default=1024)
ap.add_argument("--k", type=int, default=1000)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
fn = _lib.lib().vb_debug_hamming_bench
fn.restype = C.c_int
ms = C.c_float()
rows = (C.c_uint32 * a.k)()
dist = (C.c_float * a.k)()
rc = fn(C.c_size_t(a.rows), C.c_size_t(a.dims), C.c_size_t(a.k), a.iters, C.byref(ms), rows, dist)
assert rc == 0, rc
alg = a.rows * ((a.dims + 63) // 64) * 8
peak = 6545.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
d = list(dist)
assert d == sorted(d), "candidates not sorted by distance"
print(json.dumps({"metric": "hamming candidate scans/s (device-timed)", "value": 1e3 / ms.value, "ms_per_scan": ms.value,
                  "config": {"rows": a.rows, "dims": a.dims, "candidates": a.k},
                  "roofline": {"bound": "hbm", "achieved": alg / ms.value / 1e6, "peak": peak, "unit": "GB/s",
                               "frac": alg / ms.value / 1e6 / peak},
                  "best": [int(rows[0]), d[0]], "kth": d[-1]}))
