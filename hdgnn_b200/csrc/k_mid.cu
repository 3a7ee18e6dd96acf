// Translation unit of the fused per-commit kernel (mid2.cuh).
#include "fused.h"

namespace hdgnn {

const void* mid2_fn_rt(int cwt, bool train) {
    const void* fn = nullptr;
    HDGNN_CWT_SWITCH(cwt, fn = train ? (const void*)mid2_kernel<CWT, true> : (const void*)mid2_kernel<CWT, false>);
    return fn;
}

void launch_mid2(int cwt, bool train, int grid, size_t smem, cudaStream_t st, const Mid2Args& a) {
    if (train) { HDGNN_CWT_SWITCH(cwt, (mid2_kernel<CWT, true><<<grid, M2_T, smem, st>>>(a))); }
    else { HDGNN_CWT_SWITCH(cwt, (mid2_kernel<CWT, false><<<grid, M2_T, smem, st>>>(a))); }
}

}  // namespace hdgnn
