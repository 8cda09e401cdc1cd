// bfm_tensor.cu - launches the tensor-core matching kernel (bfm_tensor.cuh) for the library: resident inputs, k <= 2,
// no mask, no cross-check.  The scan only reduces into the row states; bfm_tiles_kernel (bfm_kernels.cuh) finalizes,
// exactly as for the static POPC form with deferred finalize, so everything after the scan is shared.
#define BFM_SCAN_INST_ONLY
#include "bfm_kernels.cuh"
#include "bfm_tensor.cuh"

namespace bfm {

constexpr int TC_STAGES = 3;   // B stages in shared memory

static_assert(sizeof(Segment) == sizeof(bfm_tc::Item), "work items of the tensor form are the planner's segments");
static_assert(offsetof(Problem, col0) == 5 * 4 && offsetof(Problem, pad) == 7 * 4 && sizeof(Problem) == 8 * 4,
              "expand_kernel reads the problem table by word offsets");
static_assert(TC_BQ == bfm_tc::BQ && TC_BT == bfm_tc::BT && TC_THREADS == bfm_tc::NTHREADS && DIST_SHIFT == bfm_tc::DIST_SHIFT_TC, "planner and kernel agree on the tile sizes");

int tensor_init() {
    return (int)cudaFuncSetAttribute(bfm_tc::scan_kernel<TC_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bfm_tc::smem_bytes(TC_STAGES));
}

int tensor_launch(const TensorLaunch &L, cudaStream_t st) {
    if (L.n_items <= 0) return 0;
    if (L.max_rows > 0) {
        const dim3 egrid((unsigned)((L.max_rows + bfm_tc::EXPAND_ROWS - 1) / bfm_tc::EXPAND_ROWS), (unsigned)L.n_problems, 2);
        // Problem: first expanded query row in `pad` (word 7), first expanded train row in `col0` (word 5)
        bfm_tc::expand_kernel<<<egrid, 256, 0, st>>>(static_cast<const uint16_t *>(L.q), static_cast<const uint16_t *>(L.t),
                                                    reinterpret_cast<const int32_t *>(L.probs), 8, 7, 5, static_cast<uint4 *>(L.xq), static_cast<uint4 *>(L.xt),
                                                    L.xq_plane / 16, L.xt_plane / 16);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    bfm_tc::Params pa;
    pa.xq = static_cast<const uint8_t *>(L.xq);
    pa.xt = static_cast<const uint8_t *>(L.xt);
    pa.xq_plane = L.xq_plane;
    pa.xt_plane = L.xt_plane;
    pa.items = reinterpret_cast<const bfm_tc::Item *>(L.items);
    pa.n_items = L.n_items;
    pa.rowstate = L.rowstate;
    pa.status = L.status;
    pa.dbg = 0;
    pa.mul_lo = (uint32_t)(-64);
    pa.mul_hi = (uint32_t)(-64) << 16;
    if (L.ev_scan[0]) cudaEventRecord(L.ev_scan[0], st);
    bfm_tc::scan_kernel<TC_STAGES><<<(unsigned)L.grid, bfm_tc::NTHREADS, bfm_tc::smem_bytes(TC_STAGES), st>>>(pa);
    if (L.ev_scan[1]) cudaEventRecord(L.ev_scan[1], st);
    return (int)cudaGetLastError();
}

}  // namespace bfm
