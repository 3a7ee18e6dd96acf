// Translation unit of the fused per-commit kernel (mid2.cuh).
#include "fused.h"

namespace hdgnn {

using Mid2Fn = void (*)(const Mid2Args);

// instantiations: hunk tables in shared memory for Nc <= 160 (CWT <= 5), in global memory for Nc > 128 (CWT >= 5); the
// training kernel exists up to 256 hunks (CWT <= 8)
template <int CWT, bool TRAIN>
static Mid2Fn mid2_pick(bool gt) {
    if constexpr (CWT <= 4) return mid2_kernel<CWT, TRAIN, false>;
    else if constexpr (CWT >= 6) return mid2_kernel<CWT, TRAIN, true>;
    else return gt ? mid2_kernel<CWT, TRAIN, true> : mid2_kernel<CWT, TRAIN, false>;
}

static Mid2Fn mid2_fn(int cwt, bool train, bool gt) {
    Mid2Fn fn = nullptr;
    // 257 .. 512 hunks: forward only, one instantiation of 16 segments (column passes and segments beyond Nc are skipped)
    if (cwt > 8) return (cwt <= 16 && gt && !train) ? mid2_kernel<16, false, true> : nullptr;
    if (!mid2_gt_supported(cwt, gt)) return nullptr;
    if (train) { HDGNN_CWT_SWITCH(cwt, fn = mid2_pick<CWT, true>(gt)); }
    else { HDGNN_CWT_SWITCH(cwt, fn = mid2_pick<CWT, false>(gt)); }
    return fn;
}

bool mid2_gt_supported(int cwt, bool gt) { return gt ? (cwt >= 5 && cwt <= 16) : (cwt >= 1 && cwt <= 5); }

const void* mid2_fn_rt(int cwt, bool train, bool gt) { return (const void*)mid2_fn(cwt, train, gt); }

void launch_mid2(int cwt, bool train, bool gt, int grid, size_t smem, cudaStream_t st, const Mid2Args& a, bool pdl) {
    Mid2Fn fn = mid2_fn(cwt, train, gt);
    if (fn) launch_ex(fn, grid, M2_T, smem, st, pdl, a);
}

}  // namespace hdgnn
