"""Self-test of the tcgen05 building blocks (TMA swizzled load, TMEM store/load, UMMA TF32 with
the 3xTF32 split) on one 128 x 32 x K tile against float64 numpy."""
import ctypes as C

import numpy as np
import pytest

from vettore_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [32, 64, 128])
@pytest.mark.parametrize("mode,tol", [(1, 2e-3), (3, 2e-6), (11, 2e-3)])
def test_tcgen05_tile_gemm(K, mode, tol):
    rng = np.random.default_rng(K + mode)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((32, K)).astype(np.float32)
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    D = np.zeros((128, 32), dtype=np.float32)
    fn = _lib.lib().vb_debug_tc_probe
    fn.restype = C.c_int
    fp = C.POINTER(C.c_float)
    rc = fn(A.ctypes.data_as(fp), B.ctypes.data_as(fp), K, mode, D.ctypes.data_as(fp))
    assert rc == 0, rc
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(D - ref).max()
    assert err <= tol, (err, D[:2, :4], ref[:2, :4])
