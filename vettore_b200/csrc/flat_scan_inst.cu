// flat_scan_inst.cu — instantiates flat_scan_kernel for ONE metric (-DVB_METRIC=<code>);
// compiled once per metric so the translation units build in parallel.
#include "flat_scan.cuh"
#include "flat_scan.h"

#ifndef VB_METRIC
#error "compile with -DVB_METRIC=<0..9>"
#endif

namespace vb {

#define VB_CAT2(a, b) a##b
#define VB_CAT(a, b) VB_CAT2(a, b)
#define VB_VARIANT(NV, R) \
    if (nv == NV && r == R) return flat_scan_kernel<VB_METRIC, NV, R>;

ScanKernel VB_CAT(flat_scan_kernel_metric_, VB_METRIC)(int nv, int r) {
    VB_VARIANT(1, 4)
    VB_VARIANT(2, 4)
    VB_VARIANT(3, 4) VB_VARIANT(3, 2)
    VB_VARIANT(4, 4) VB_VARIANT(4, 2)
    VB_VARIANT(6, 2) VB_VARIANT(6, 1) VB_VARIANT(6, 4)
    VB_VARIANT(8, 2) VB_VARIANT(8, 1)
    VB_VARIANT(12, 1)
    VB_VARIANT(0, 2)
    return nullptr;
}

}  // namespace vb
