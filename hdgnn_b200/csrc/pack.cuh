// pack: dense uint8 label grids -> label bitmaps (bits.cuh).  Included by hdgnn.cu only.
#pragma once
#include "bits.cuh"

namespace hdgnn {

// one warp per row of either grid
__global__ void __launch_bounds__(256) pack_bits_kernel(const PackArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long rows_e = (long long)a.B * a.Ne, rows_c = (long long)a.B * a.Nc;
    const uint8_t* src; uint32_t* dst; int N, WP, r;
    if (row < rows_e) {
        r = (int)(row % a.Ne); N = a.Ne; WP = a.WPe;
        src = a.adj + (size_t)row * a.pe; dst = a.ebits + (size_t)row * WP;
    } else if (row < rows_e + rows_c) {
        const long long rr = row - rows_e;
        r = (int)(rr % a.Nc); N = a.Nc; WP = a.WPc;
        src = a.Y + (size_t)rr * a.pc; dst = a.ybits + (size_t)rr * WP;
    } else {
        return;
    }
    uint32_t mine = 0u;
    const int cw = (N + 31) >> 5;
    for (int sg = 0; sg < cw; ++sg) {
        const int c = sg * 32 + lane;
        const bool v = c < N && c != r && src[c] != 0;
        const uint32_t w = __ballot_sync(0xffffffffu, v);
        if (lane == sg) mine = w;
    }
    if (lane < WP) dst[lane] = mine;
}

}  // namespace hdgnn
