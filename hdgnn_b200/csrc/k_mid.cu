// Translation unit of the fused per-commit kernel (mid2.cuh).
#include "fused.h"

namespace hdgnn {

using Mid2Fn = void (*)(const Mid2Args);

// instantiations: hunk tables in shared memory for Nc <= 160 (CWT <= 5), in global memory for Nc > 128 (CWT >= 5); the
// training kernel exists up to 256 hunks (CWT <= 8); the cluster form (CL) for the shared-memory tables up to 128 hunks
template <int CWT, bool TRAIN>
static Mid2Fn mid2_pick(bool gt, bool cl) {
    if constexpr (CWT <= 4) return cl ? mid2_kernel<CWT, TRAIN, false, true> : mid2_kernel<CWT, TRAIN, false, false>;
    else if constexpr (CWT >= 6) return mid2_kernel<CWT, TRAIN, true, false>;
    else return gt ? mid2_kernel<CWT, TRAIN, true, false> : mid2_kernel<CWT, TRAIN, false, false>;
}

static Mid2Fn mid2_fn(int cwt, bool train, bool gt, bool cl) {
    Mid2Fn fn = nullptr;
    if (cl && (gt || cwt > 4)) return nullptr;
    // 257 .. 512 hunks: forward only, one instantiation of 16 segments (column passes and segments beyond Nc are skipped)
    if (cwt > 8) return (cwt <= 16 && gt && !train) ? mid2_kernel<16, false, true, false> : nullptr;
    if (!mid2_gt_supported(cwt, gt)) return nullptr;
    if (train) { HDGNN_CWT_SWITCH(cwt, fn = mid2_pick<CWT, true>(gt, cl)); }
    else { HDGNN_CWT_SWITCH(cwt, fn = mid2_pick<CWT, false>(gt, cl)); }
    return fn;
}

bool mid2_gt_supported(int cwt, bool gt) { return gt ? (cwt >= 5 && cwt <= 16) : (cwt >= 1 && cwt <= 5); }

const void* mid2_fn_rt(int cwt, bool train, bool gt, bool cl) { return (const void*)mid2_fn(cwt, train, gt, cl); }

void launch_mid2(int cwt, bool train, bool gt, bool cl, int grid, size_t smem, cudaStream_t st, const Mid2Args& a, bool pdl) {
    Mid2Fn fn = mid2_fn(cwt, train, gt, cl);
    if (fn) launch_ex(fn, grid, M2_T, smem, st, pdl, a, cl ? 2 : 1);
}

// co-resident clusters of two CTAs of the cluster form at this shared-memory size (0: not available)
int mid2_max_clusters(int cwt, bool train, size_t smem) {
    Mid2Fn fn = mid2_fn(cwt, train, false, true);
    if (!fn) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(M2_T); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)fn, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // namespace hdgnn
