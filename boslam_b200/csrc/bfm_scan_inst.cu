// bfm_scan_inst.cu - instantiates the matching kernel (bfm_kernels.cuh) for ONE register tile and mode.
// Compiled nine times (INST_R in {1, 2, 4} x INST_MODE in {0: k = 1, 1: cross-check, 2: k = 2}) so the variants
// build in parallel; bfm_api.cu picks one through bfm_pick_scan_r<R>_m<MODE>().
#define BFM_SCAN_INST_ONLY
#include "bfm_kernels.cuh"

#ifndef INST_R
#error "compile with -DINST_R=1|2|4 -DINST_MODE=0|1|2"
#endif

namespace bfm {

constexpr int INST_NT = 128;

template <int MASK, int PM>
static ScanFn inst_fn(bool dyn) {
    if (dyn) return bfm_scan_persistent_kernel<INST_R, (INST_MODE == 2 ? 2 : 1), (INST_MODE == 1), MASK, PM, INST_NT, false>;
    return bfm_scan_static_kernel<INST_R, (INST_MODE == 2 ? 2 : 1), (INST_MODE == 1), MASK, PM, INST_NT, false>;
}
template <int MASK>
static ScanFn inst_pm(int pm, bool dyn) {
    switch (pm) {
        case 4: return inst_fn<MASK, 4>(dyn);
        case 5: return inst_fn<MASK, 5>(dyn);
        case 6: return inst_fn<MASK, 6>(dyn);
        case 40: return inst_fn<MASK, 40>(dyn);
        case 50: return inst_fn<MASK, 50>(dyn);
        default: return inst_fn<MASK, 8>(dyn);
    }
}

#define BFM_CAT2(a, b, c, d) a##b##c##d
#define BFM_CAT(a, b, c, d) BFM_CAT2(a, b, c, d)

// mask: 0 none, 1 dense, 2 window; bound: the k > 2 pass variant (only R = 1, MODE = 2 has it)
ScanFn BFM_CAT(pick_scan_r, INST_R, _m, INST_MODE)(int mask, int pm, bool bound, bool dyn) {
#if INST_R == 1 && INST_MODE == 2
    if (bound) {
        switch (mask * 2 + (dyn ? 1 : 0)) {
            case 2: return bfm_scan_static_kernel<1, 2, false, 1, 40, INST_NT, true>;
            case 3: return bfm_scan_persistent_kernel<1, 2, false, 1, 40, INST_NT, true>;
            case 4: return bfm_scan_static_kernel<1, 2, false, 2, 40, INST_NT, true>;
            case 5: return bfm_scan_persistent_kernel<1, 2, false, 2, 40, INST_NT, true>;
            case 1: return bfm_scan_persistent_kernel<1, 2, false, 0, 40, INST_NT, true>;
            default: return bfm_scan_static_kernel<1, 2, false, 0, 40, INST_NT, true>;
        }
    }
#else
    (void)bound;
#endif
    switch (mask) {
        case 1: return inst_pm<1>(pm, dyn);
        case 2: return inst_pm<2>(pm, dyn);
        default: return inst_pm<0>(pm, dyn);
    }
}

}  // namespace bfm
