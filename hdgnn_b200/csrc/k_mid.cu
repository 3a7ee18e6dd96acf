// Translation unit of the fused per-commit kernel (mid2.cuh).
#include "fused.h"

namespace hdgnn {

const void* mid2_fn_rt(int cwt, bool train) {
    const void* fn = nullptr;
    HDGNN_CWT_SWITCH(cwt, fn = train ? (const void*)mid2_kernel<CWT, true> : (const void*)mid2_kernel<CWT, false>);
    return fn;
}

void launch_mid2(int cwt, bool train, int grid, size_t smem, cudaStream_t st, const Mid2Args& a, bool pdl) {
    if (train) { HDGNN_CWT_SWITCH(cwt, launch_ex(mid2_kernel<CWT, true>, grid, M2_T, smem, st, pdl, a)); }
    else { HDGNN_CWT_SWITCH(cwt, launch_ex(mid2_kernel<CWT, false>, grid, M2_T, smem, st, pdl, a)); }
}

}  // namespace hdgnn
