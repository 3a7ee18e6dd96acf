// sweep: the O(N^2 * 20) inner loops of the hot path, shared by the entity-grid kernels (ent.cuh)
// and the fused per-commit kernel (mid.cuh).
//
// Every per-pair layer of the reference has the separable first-layer form
//        pre_ij[k] = P_i[k] + Q_j[k] + l_ij * D[k],      l_ij in {0,1}
// (marshalling + 4->20 / 10->20 / 22->20 matmul, model_2.py:144,170,260,315), and its output is
// only ever consumed through row sums over j, column sums over i, or the 20->2 head.
//
// Mapping (one warp = a group of rows, lanes = 32 consecutive columns, channels k in registers):
//   * Q_j[k] and the column accumulators live in registers for a whole column block;
//   * the label enters through the ADDRESS, not the arithmetic: each row keeps two tables
//     P0_i[k] = P_i[k] and P1_i[k] = P_i[k] + D[k]; a lane loads its 20 values from P0 or P1
//     according to its own label byte (two smem wavefronts instead of a broadcast), so the
//     pre-activation costs ONE FADD per (pair, channel);
//   * row sums: the 20 per-lane values are transpose-reduced across the warp (21 shuffles per
//     32 pairs) and accumulated, in column-block order, into a per-row table by the one warp
//     that owns the row;  column sums: per-warp partials are combined through shared memory in
//     warp order.  No atomics anywhere => bitwise run-to-run determinism.
//   * the diagonal pair (i,i) is swept with l = 0 and subtracted afterwards by the callers.
#pragma once
#include "common.cuh"

namespace hdgnn {

// Which index of a 20-vector this lane holds after warp_reduce20, or -1.
__device__ __forceinline__ int reduce20_index(int lane) {
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
    if (b0) return (b1 | b2) ? -1 : b4 * 10 + b3 * 5 + 4;
    return b4 * 10 + b3 * 5 + (b1 ? (b2 ? 3 : 1) : (b2 ? 2 : 0));
}

// Sum each of v[0..19] over the 32 lanes.  Afterwards lane L holds the total of index
// reduce20_index(L).  Fixed exchange order.
__device__ __forceinline__ float warp_reduce20(const float (&v)[HD], int lane) {
    float a[10];
    const bool h16 = lane & 16;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float keep = h16 ? v[i + 10] : v[i], send = h16 ? v[i] : v[i + 10];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float b[5];
    const bool h8 = lane & 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float keep = h8 ? a[i + 5] : a[i], send = h8 ? a[i] : a[i + 5];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const bool h4 = lane & 4;
    float c0, c1, c2;
    {
        float keep = h4 ? b[2] : b[0], send = h4 ? b[0] : b[2];
        c0 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        keep = h4 ? b[3] : b[1]; send = h4 ? b[1] : b[3];
        c1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        c2 = b[4] + __shfl_xor_sync(0xffffffffu, b[4], 4);
    }
    const bool h2 = lane & 2;
    float d0, d1;
    {
        const float keep = h2 ? c1 : c0, send = h2 ? c0 : c1;
        d0 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        d1 = c2 + __shfl_xor_sync(0xffffffffu, c2, 2);
    }
    const bool h1 = lane & 1;
    const float keep = h1 ? d1 : d0, send = h1 ? d0 : d1;
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ void load20(float (&dst)[HD], const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const float4 t = s4[q];
        dst[4 * q] = t.x; dst[4 * q + 1] = t.y; dst[4 * q + 2] = t.z; dst[4 * q + 3] = t.w;
    }
}

// Shared-memory scratch for combining per-warp column partials: [warp][k][lane].
__device__ __forceinline__ void store_col_partial(float* cpart, int warp, int lane, const float (&col)[HD]) {
    float* dst = cpart + (size_t)warp * (HD * 32) + lane;
#pragma unroll
    for (int k = 0; k < HD; ++k) dst[k * 32] = col[k];
}
// Sum over warps (fixed order) of entry (k, ln); call after a __syncthreads().
__device__ __forceinline__ float combine_col_partial(const float* cpart, int nwarps, int k, int ln) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += cpart[(size_t)w * (HD * 32) + k * 32 + ln];
    return acc;
}

// ---- forward pair sum ------------------------------------------------------------------------
// One warp, one column block: rows r = row_first, row_first + row_step, ... < nrows (tile-local).
//   P01     [2][rows_cap][20]  tables (P0 then P1), shared memory
//   labrow0 pointer to the label byte of (tile row 0, this lane's column); row stride `pitch`
//   grow0   global row index of tile row 0 (to zero the diagonal label)
//   col_ok  this lane's column is < N
//   Q       this lane's column values;  col  column accumulators (in/out)
//   rowacc  [rows_cap][20] shared: per-row sums, accumulated across column blocks by this warp
template <typename LabT>
__device__ __forceinline__ void sweep_fwd_block(const float* P01, int rows_cap, const LabT* labrow0, int pitch,
                                                int grow0, int gcol, bool col_ok, int nrows, int row_first,
                                                int row_step, const float (&Q)[HD], float (&col)[HD], float* rowacc,
                                                int lane, int ridx) {
    for (int r = row_first; r < nrows; r += row_step) {
        const bool lab = col_ok && (gcol != grow0 + r) && (labrow0[(size_t)r * pitch] != 0);
        float P[HD];
        load20(P, P01 + ((size_t)(lab ? rows_cap : 0) + r) * HD);
        float h[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            h[k] = fmaxf(P[k] + Q[k], 0.f);
            col[k] += h[k];
        }
        const float tot = warp_reduce20(h, lane);
        if (ridx >= 0) rowacc[r * HD + ridx] += tot;
    }
}

// ---- backward pair sum -----------------------------------------------------------------------
//   v_ij[k] = [pre_ij[k] > 0] * (GR_i[k] + GC_j[k])
//   rowacc += sum_j v ; col += sum_i v ; lacc += sum_{l_ij = 1} v
template <typename LabT>
__device__ __forceinline__ void sweep_bwd_block(const float* P01, const float* GRt, int rows_cap, const LabT* labrow0,
                                                int pitch, int grow0, int gcol, bool col_ok, int nrows, int row_first,
                                                int row_step, const float (&Q)[HD], const float (&GC)[HD],
                                                float (&col)[HD], float (&lacc)[HD], float* rowacc, int lane, int ridx) {
    for (int r = row_first; r < nrows; r += row_step) {
        const bool lab = col_ok && (gcol != grow0 + r) && (labrow0[(size_t)r * pitch] != 0);
        float P[HD], G[HD];
        load20(P, P01 + ((size_t)(lab ? rows_cap : 0) + r) * HD);
        load20(G, GRt + (size_t)r * HD);
        float v[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            const float t = P[k] + Q[k];
            const float g = G[k] + GC[k];
            v[k] = t > 0.f ? g : 0.f;
            col[k] += v[k];
            if (lab) lacc[k] += v[k];
        }
        const float tot = warp_reduce20(v, lane);
        if (ridx >= 0) rowacc[r * HD + ridx] += tot;
    }
}

}  // namespace hdgnn
