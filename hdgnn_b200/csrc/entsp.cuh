// entsp: the entity pair layer (model_2.py:161-188: per-pair MLP 4 -> 20 (relu) -> 20 and its row / column
// aggregation) and its backward WITHOUT visiting the Ne x Ne grid.
//
// The first layer of the pair MLP is rank-1 in the node attribute:  pre_ij[k] = U_k x_i + V_k x_j + c_k + l_ij D_k
// (c = b1 + W_label[0], D = W_label[1] - W_label[0]; model_2.py:144,167-170), and only its row and column sums are
// consumed.  Split   relu(pre_ij) = relu(pre0_ij) + l_ij (relu(pre0_ij + D) - relu(pre0_ij)),  pre0 = the l = 0 value:
//   * dense part: for a fixed (i, k) the sign of pre0_ij is monotone in x_j, so over the nodes SORTED by x the active
//     set {j : pre0_ij[k] > 0} is a suffix (V_k >= 0) or a prefix (V_k < 0) found by one binary search, and
//         sum_j relu(pre0_ij[k]) = cnt (U_k x_i + c_k) + V_k * (sum of x over the active set)
//     comes from a suffix-sum table: O(Ne log Ne) per channel instead of O(Ne^2).  The predicate evaluated by the
//     search is the same fused multiply-add in the forward and in the backward, so the relu gates agree bit for bit;
//   * sparse part: one walk over the set bits of the label bitmap (rows for the sources, the transposed bitmap for
//     the targets) adds the correction of the l = 1 pairs: O(nnz) per channel.
// The backward has the same structure: v_ij[k] = [pre_ij[k] > 0] (G_i[k] + G_j[k]) summed over an active prefix /
// suffix needs cnt * G_i[k] + (suffix sums of G in sorted order), and the l = 1 pairs are corrected edge by edge.
// Everything is accumulated by ONE owner thread per (node, channel pair) in a fixed order: no atomics, bitwise
// run-to-run determinism.  Work per commit at glide (Ne = 200, 5 % density): ~0.3 M lane operations instead of the
// 4 M of the dense sweep; at 50 % density the edge walk costs more than the dense sweep (ent2.cuh), which remains
// selectable (HDGNN_F_DENSE_SWEEP).
#pragma once
#include "common.cuh"

namespace hdgnn {

// first rank r in [0, N] whose node is on the far side of the threshold: with g(r) = (w >= 0) == (fma(xsort[r], w, base) > 0),
// g is monotone (false ... false true ... true) over the ranks; P2 = smallest power of two > N.
__device__ __forceinline__ int entsp_split(const float* xsort, int N, int P2, float w, float base) {
    const bool pos = w >= 0.f;
    int r = 0;
    for (int step = P2 >> 1; step > 0; step >>= 1) {
        const int probe = r + step - 1;
        const float xv = xsort[probe < N ? probe : N - 1];
        const bool act = fmaf(xv, w, base) > 0.f;
        if (probe < N && act != pos) r += step;
    }
    return r;
}

// number of active nodes and the sum of a sorted-order suffix table over them: suf[r] = sum_{r' >= r} t[r'], suf[N] = 0
__device__ __forceinline__ void entsp_active(int r, int N, bool pos, const float* suf, float& cnt, float& sum) {
    cnt = (float)(pos ? N - r : r);
    sum = pos ? suf[r] : suf[0] - suf[r];
}

// 32 x 32 bit-matrix transpose across a warp: lane a holds row a; afterwards lane b holds column b (bit a = old row a, bit b)
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
        m ^= m << (j >> 1);
    }
    return x;
}

// in-place exclusive prefix sum of p[0..n) by ONE warp; p[n] receives the total
__device__ __forceinline__ void warp_excl_scan_int(int* p, int n, int lane) {
    const int per = (n + 31) >> 5, lo = min(lane * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += p[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int run = incl - sum;
    for (int i = lo; i < hi; ++i) { const int v = p[i]; p[i] = run; run += v; }
    if (lane == 31) p[n] = incl;
}

// Cursor over the set bits of an n-row bitmap in row-major order, positioned at the edge with index e0 given the
// exclusive prefix counts ptr[0..n] (ptr[n] = number of edges > e0).  P2 = smallest power of two > n.
struct EdgeCursor { int row, w; uint32_t bits; };
__device__ __forceinline__ EdgeCursor edge_seek(const uint32_t* bm, int WP, const int* ptr, int n, int P2, int e0) {
    int row = 0;                                           // largest row with ptr[row] <= e0 (that row holds edge e0)
    for (int step = P2 >> 1; step > 0; step >>= 1) {
        const int probe = row + step;
        if (probe < n && ptr[probe] <= e0) row = probe;
    }
    int skip = e0 - ptr[row], w = 0;
    uint32_t bits = bm[(size_t)row * WP];
    for (int c = __popc(bits); skip >= c; c = __popc(bits)) { skip -= c; ++w; bits = bm[(size_t)row * WP + w]; }
    for (; skip > 0; --skip) bits &= bits - 1;
    return {row, w, bits};
}

}  // namespace hdgnn
