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
    VB_VARIANT(1, 4) VB_VARIANT(1, 8)
    VB_VARIANT(2, 4)
    VB_VARIANT(3, 4) VB_VARIANT(3, 2)
    VB_VARIANT(4, 4) VB_VARIANT(4, 2)
    VB_VARIANT(6, 2) VB_VARIANT(6, 1) VB_VARIANT(6, 4)
    VB_VARIANT(8, 2) VB_VARIANT(8, 1)
    VB_VARIANT(12, 1)
    VB_VARIANT(0, 2)
    return nullptr;
}

#define VB_STREAM_VARIANT(NV, RPW, W) \
    if (nv == NV && rpw == RPW && warps == W) return flat_stream_kernel<VB_METRIC, NV, RPW, W>;

StreamKernel VB_CAT(flat_stream_kernel_metric_, VB_METRIC)(int nv, int rpw, int warps) {
    VB_STREAM_VARIANT(1, 4, 16)
    VB_STREAM_VARIANT(2, 4, 16)
    VB_STREAM_VARIANT(3, 2, 16)
    VB_STREAM_VARIANT(4, 2, 16) VB_STREAM_VARIANT(4, 1, 16)
    VB_STREAM_VARIANT(6, 1, 16) VB_STREAM_VARIANT(6, 2, 8) VB_STREAM_VARIANT(6, 1, 8) VB_STREAM_VARIANT(6, 2, 16)
    VB_STREAM_VARIANT(8, 1, 16)
    VB_STREAM_VARIANT(12, 1, 8)
    return nullptr;
}

}  // namespace vb
