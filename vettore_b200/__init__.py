"""vettore_b200 — B200-native (sm_100a) scan path of Vettore behind the reference's own
NIF surface.

* ``vettore_b200.nifs``  — function-for-function mirror of ``Vettore.Nifs``
  (reference lib/vettore_nifs.ex) over the C ABI in ``include/vettore_b200.h``.
* ``vettore_b200.index`` — ``Vettore.Index.Flat``-shaped adapter (reference
  lib/vettore/index/flat.ex) plus the search pipelines of ``Vettore.Collection``
  that sit on the scan path.
* ``vettore_b200.sharded`` — row-sharded multi-GPU search (one process per GPU,
  torch.distributed all-gather of the per-GPU top-k + K7 merge).

There is no CPU fallback: importing works anywhere, every compute call needs a CUDA
device and ``libvettore_b200.so`` (built by ``vettore_b200/build.py`` /
``__graft_entry__.build``).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
