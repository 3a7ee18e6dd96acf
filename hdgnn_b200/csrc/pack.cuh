// pack: dense uint8 label grids -> label bitmaps (bits.cuh).  Included by hdgnn.cu only.
// A row (pitch <= 512 bytes, multiple of 16) is read with one 16-byte load per thread by a group of
// TPR = 16 or 32 threads; each thread turns its 16 bytes into a 16-bit mask and even threads combine
// two masks into one 32-column word.
#pragma once
#include "bits.cuh"

namespace hdgnn {

__device__ __forceinline__ uint32_t nz_mask4(uint32_t v) {        // bit i = byte i of v != 0
    const uint32_t t = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
    return ((t >> 7) | (t >> 14) | (t >> 21) | (t >> 28)) & 0xfu;
}

template <int TPR>      // threads per row: 16 (pitch <= 256) or 32
__global__ void __launch_bounds__(256) pack_bits_kernel(const PackArgs a) {
    const int lane = threadIdx.x & 31, sub = threadIdx.x & (TPR - 1);
    const long long row = (long long)blockIdx.x * (256 / TPR) + (threadIdx.x / TPR);
    const long long rows_e = (long long)a.B * a.Ne, rows_c = (long long)a.B * a.Nc;
    const uint8_t* src = nullptr; uint32_t* dst = nullptr; int N = 0, WP = 0, r = 0, pitch = 0;
    if (row < rows_e) {
        r = (int)(row % a.Ne); N = a.Ne; WP = a.WPe; pitch = a.pe;
        src = a.adj + (size_t)row * a.pe; dst = a.ebits + (size_t)row * WP;
    } else if (row < rows_e + rows_c) {
        const long long rr = row - rows_e;
        r = (int)(rr % a.Nc); N = a.Nc; WP = a.WPc; pitch = a.pc;
        src = a.Y + (size_t)rr * a.pc; dst = a.ybits + (size_t)rr * WP;
    }
    uint32_t m = 0u;
    if (src && sub * 16 < pitch) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + sub * 16);
        m = nz_mask4(v.x) | (nz_mask4(v.y) << 4) | (nz_mask4(v.z) << 8) | (nz_mask4(v.w) << 12);
    }
    const uint32_t hi = __shfl_down_sync(0xffffffffu, m, 1);
    if (src && (sub & 1) == 0) {
        const int sg = sub >> 1;
        if (sg < WP) {
            uint32_t w = m | (hi << 16);
            const int c0 = sg * 32;
            if (c0 + 32 > N) w &= (c0 >= N) ? 0u : (0xffffffffu >> (c0 + 32 - N));     // columns >= N
            if ((r >> 5) == sg) w &= ~(1u << (r & 31));                                   // diagonal
            dst[sg] = w;
        }
    }
    (void)lane;
    // Launched with programmatic serialization right after the previous step's optimizer kernel: the packing
    // above only reads this step's inputs and writes this step's bitmap buffer (two buffers alternate), so it
    // overlaps that kernel; completing only after it has finished keeps every later kernel of this step ordered.
    pdl_wait();
}

}  // namespace hdgnn
