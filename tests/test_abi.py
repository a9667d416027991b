"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly the
symbols include/vettore_b200.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vettore_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_nif_surface_for_the_scan_path():
    syms = declared_symbols()
    for needed in ["vb_flat_new", "vb_flat_insert", "vb_flat_insert_many", "vb_flat_delete", "vb_flat_search",
                   "vb_vector_top_k", "vb_binary_top_k", "vb_compress_sign_bits", "vb_last_error"]:
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from vettore_b200 import _lib
    from vettore_b200.build import build
    build()
    L = ctypes.CDLL(_lib.SO_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    from vettore_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_device_is_a_loud_error_not_a_fallback():
    from vettore_b200 import _lib, nifs
    if _lib.lib().vb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="cuda"):
        nifs.flat_new_cosine()
    res = nifs.vector_top_k([("a", [1.0, 0.0])], [1.0, 0.0], 0, 2, 1)
    assert res[0] == "error" and res[1].startswith("cuda:")


def test_validation_errors_do_not_need_a_device():
    """Reference error strings that are raised before any row is scored (search.rs:45-48, :84)."""
    from vettore_b200 import nifs
    assert nifs.vector_top_k([], [1.0], 9, 1, 1) == ("error", "unknown metric")
    assert nifs.vector_top_k([], [1.0], 0, 0, 1) == ("error", "invalid prefix dimensions")
    assert nifs.vector_top_k([], [1.0], 0, 2, 1) == ("error", "invalid prefix dimensions")
    assert nifs.vector_top_k([], [float("nan")], 0, 1, 1) == ("error", "vector contains a non-finite value")
    assert nifs.vector_top_k([], [1.0], 0, 1, 1) == ("ok", [])
    assert nifs.binary_top_k([], [], 0, 1) == ("error", "dimensions must be positive")
    assert nifs.binary_top_k([], [], 1, 1) == ("error", "dimension mismatch")
    assert nifs.binary_top_k([], [0], 1, 1) == ("ok", [])
    assert nifs.compress_sign_bits([1.0, -1.0, 0.0]) == [5]
    assert nifs.compress_sign_bits([-0.0]) == [1]


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vettore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
