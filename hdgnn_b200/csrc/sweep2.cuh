// sweep2: the O(N^2 * 20) inner loops of the hot path (entity grid: model_2.py:161-188; hunk grid:
// model_2.py:245-277, 304-324) and their backward, shared by ent2.cuh and mid2.cuh.
//
// Every per-pair layer of the reference has the separable first-layer form
//        pre_ij[k] = P_i[k] + Q_j[k] + l_ij * D[k],      l_ij in {0,1}
// and its output is only consumed through row sums over j and column sums over i.
//
// Mapping ("channel-group warps"):
//   * a warp owns ONE group of 4 channels (kg = 0..4) and a subset of the rows (rg = 0..NRG-1,
//     rows rg, rg+NRG, ...); its lanes own the columns sg*32 + lane, sg < CW, ALL at once, so the
//     row sum of a row is accumulated in-lane over the CW segments and crosses lanes only once per
//     row (one 4-value transpose-reduce), and the column sums never leave the lane;
//   * Q_j and the column accumulators live in registers as packed f32x2 pairs; the arithmetic that
//     can be packed (pre-activation add, both accumulations) uses the sm_100 f32x2 add / fma;
//   * labels are a bitmap (bits.cuh): word sg of a row holds l_{i, sg*32 + lane} at bit `lane`, the
//     diagonal bit is zero.  The label enters through the ADDRESS: each row keeps
//     P01[r][kg][0] = P_i, P01[r][kg][1] = P_i + D (4 floats each) and a lane loads 16 bytes from
//     one or the other, so the pre-activation costs one packed add per two channels;
//   * the diagonal pair (i,i) is swept with l = 0 and subtracted afterwards by the callers;
//   * no atomics: every sum has a fixed order => bitwise run-to-run determinism.
#pragma once
#include "common.cuh"

namespace hdgnn {

constexpr int KG = 5;                 // channel groups of 4 (HD = 20)
constexpr int PROW = KG * 8;          // floats per row of a P01 table: [kg][2][4]

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 relu2(u64 v) {
    float lo, hi;
    upk2(v, lo, hi);
    return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}
// (t > 0) ? g : 0 per half
__device__ __forceinline__ u64 gate2(u64 t, u64 g) {
    float tl, th, gl, gh;
    upk2(t, tl, th);
    upk2(g, gl, gh);
    return pk2(tl > 0.f ? gl : 0.f, th > 0.f ? gh : 0.f);
}

// Sum 4 per-lane values (two packed pairs = channels 0..3 of the warp's group) over the 32 lanes.
// Afterwards lanes 0, 8, 16, 24 hold the totals of channels 0, 2, 1, 3 (see reduce4_channel).
// Fixed exchange order.
__device__ __forceinline__ float reduce4(u64 a01, u64 a23, int lane) {
    float a0, a1, a2, a3;
    upk2(a01, a0, a1);
    upk2(a23, a2, a3);
    const bool h16 = lane & 16;
    float k0 = h16 ? a1 : a0, s0 = h16 ? a0 : a1;
    float k1 = h16 ? a3 : a2, s1 = h16 ? a2 : a3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool h8 = lane & 8;
    float k = h8 ? k1 : k0;
    const float s = h8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}
// channel (0..3) whose total a lane with (lane & 7) == 0 holds after reduce4
__device__ __forceinline__ int reduce4_channel(int lane) { return ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1); }

// label words of one row -> registers (row-uniform 16-byte loads)
template <int CW>
__device__ __forceinline__ void load_words(uint32_t (&w)[CW], const uint32_t* row) {
    constexpr int NQ = (CW + 3) / 4;
    uint32_t t[NQ * 4];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const uint4 v = reinterpret_cast<const uint4*>(row)[q];
        t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int s = 0; s < CW; ++s) w[s] = t[s];
}

// Sum 8 per-lane values (4 channels of two rows: a = row r, c = row r2) over the 32 lanes.  Afterwards the
// lanes with (lane & 3) == 0 hold the total of row ((lane >> 4) & 1), channel reduce8_channel(lane).
__device__ __forceinline__ float reduce8(u64 a01, u64 a23, u64 c01, u64 c23, int lane) {
    float a0, a1, a2, a3, c0, c1, c2, c3;
    upk2(a01, a0, a1); upk2(a23, a2, a3); upk2(c01, c0, c1); upk2(c23, c2, c3);
    const bool h16 = lane & 16;
    float k0 = h16 ? c0 : a0, k1 = h16 ? c1 : a1, k2 = h16 ? c2 : a2, k3 = h16 ? c3 : a3;
    const float s0 = h16 ? a0 : c0, s1 = h16 ? a1 : c1, s2 = h16 ? a2 : c2, s3 = h16 ? a3 : c3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16); k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    k2 += __shfl_xor_sync(0xffffffffu, s2, 16); k3 += __shfl_xor_sync(0xffffffffu, s3, 16);
    const bool h8 = lane & 8;
    float m0 = h8 ? k2 : k0, m1 = h8 ? k3 : k1;
    const float t0 = h8 ? k0 : k2, t1 = h8 ? k1 : k3;
    m0 += __shfl_xor_sync(0xffffffffu, t0, 8); m1 += __shfl_xor_sync(0xffffffffu, t1, 8);
    const bool h4 = lane & 4;
    float k = h4 ? m1 : m0;
    const float t = h4 ? m0 : m1;
    k += __shfl_xor_sync(0xffffffffu, t, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}
__device__ __forceinline__ int reduce8_channel(int lane) { return 2 * ((lane >> 3) & 1) + ((lane >> 2) & 1); }

// ---- forward pair sum ------------------------------------------------------------------------
//   P01   smem [rows][KG][2][4], 32-byte aligned       bits  smem [rows][WP], this pass's words first
//   Q     this lane's column values per segment (packed), NEG_BIG in both halves for columns >= N
//   col   column accumulators (in/out)                  rowacc smem [rows][20]
// ACCUM: add to rowacc instead of storing (later column passes of a wide grid).
// One row of one warp: per-lane partial row sums rp0/rp1 (channels 0,1 / 2,3), column sums updated.
template <int CW>
__device__ __forceinline__ void sweep2_fwd_row(const float* P01, const uint32_t* bits, int WP, int r, int kg, uint32_t lmask,
                                               const u64 (&Q)[CW][2], u64 (&col)[CW][2], u64& rp0, u64& rp1) {
    uint32_t w[CW];
    load_words<CW>(w, bits + (size_t)r * WP);
    const float* prow0 = P01 + (size_t)r * PROW + kg * 8;
    const float* prow1 = prow0 + 4;
    rp0 = 0ull; rp1 = 0ull;
#pragma unroll
    for (int sg = 0; sg < CW; ++sg) {
        const ulonglong2 p = *reinterpret_cast<const ulonglong2*>((w[sg] & lmask) ? prow1 : prow0);
        const u64 h0 = relu2(add2(p.x, Q[sg][0])), h1 = relu2(add2(p.y, Q[sg][1]));
        col[sg][0] = add2(col[sg][0], h0); col[sg][1] = add2(col[sg][1], h1);
        rp0 = add2(rp0, h0); rp1 = add2(rp1, h1);
    }
}

template <bool ACCUM>
__device__ __forceinline__ void row_store(float* rowacc, int r, int k, float tot) {
    float* dst = rowacc + (size_t)r * HD + k;
    if (ACCUM) *dst += tot; else *dst = tot;
}

// rows are taken two at a time so that one 8-value transpose-reduce serves both
template <int CW, bool ACCUM>
__device__ __forceinline__ void sweep2_fwd(const float* P01, const uint32_t* bits, int WP, int nrows, int rg, int nrg,
                                           int kg, const u64 (&Q)[CW][2], u64 (&col)[CW][2], float* rowacc, int lane) {
    const uint32_t lmask = 1u << lane;
    int r = rg;
    // four rows per trip: the two transpose-reduces are independent, so one hides behind the other's shuffle latency
    for (; r + 3 * nrg < nrows; r += 4 * nrg) {
        u64 a0, a1, c0, c1, e0, e1, g0, g1;
        sweep2_fwd_row<CW>(P01, bits, WP, r, kg, lmask, Q, col, a0, a1);
        sweep2_fwd_row<CW>(P01, bits, WP, r + nrg, kg, lmask, Q, col, c0, c1);
        sweep2_fwd_row<CW>(P01, bits, WP, r + 2 * nrg, kg, lmask, Q, col, e0, e1);
        sweep2_fwd_row<CW>(P01, bits, WP, r + 3 * nrg, kg, lmask, Q, col, g0, g1);
        const float tot = reduce8(a0, a1, c0, c1, lane), tot2 = reduce8(e0, e1, g0, g1, lane);
        if ((lane & 3) == 0) {
            row_store<ACCUM>(rowacc, (lane & 16) ? r + nrg : r, kg * 4 + reduce8_channel(lane), tot);
            row_store<ACCUM>(rowacc, (lane & 16) ? r + 3 * nrg : r + 2 * nrg, kg * 4 + reduce8_channel(lane), tot2);
        }
    }
    for (; r + nrg < nrows; r += 2 * nrg) {
        u64 a0, a1, c0, c1;
        sweep2_fwd_row<CW>(P01, bits, WP, r, kg, lmask, Q, col, a0, a1);
        sweep2_fwd_row<CW>(P01, bits, WP, r + nrg, kg, lmask, Q, col, c0, c1);
        const float tot = reduce8(a0, a1, c0, c1, lane);
        if ((lane & 3) == 0) row_store<ACCUM>(rowacc, (lane & 16) ? r + nrg : r, kg * 4 + reduce8_channel(lane), tot);
    }
    if (r < nrows) {
        u64 a0, a1;
        sweep2_fwd_row<CW>(P01, bits, WP, r, kg, lmask, Q, col, a0, a1);
        const float tot = reduce4(a0, a1, lane);
        if ((lane & 7) == 0) row_store<ACCUM>(rowacc, r, kg * 4 + reduce4_channel(lane), tot);
    }
}

// ---- backward pair sum -----------------------------------------------------------------------
//   v_ij[k] = [pre_ij[k] > 0] * (GR_i[k] + GC_j[k])
//   rowacc = sum_j v ; col += sum_i v ; lacc += sum_{l_ij = 1} v
//   GRt   smem [rows][20]
template <int CW>
__device__ __forceinline__ void sweep2_bwd_row(const float* P01, const float* GRt, const uint32_t* bits, int WP, int r, int kg,
                                               uint32_t lmask, const u64 (&Q)[CW][2], const u64 (&GC)[CW][2],
                                               u64 (&col)[CW][2], u64 (&lacc)[2], u64& rp0, u64& rp1) {
    uint32_t w[CW];
    load_words<CW>(w, bits + (size_t)r * WP);
    const float* prow0 = P01 + (size_t)r * PROW + kg * 8;
    const float* prow1 = prow0 + 4;
    const ulonglong2 g = *reinterpret_cast<const ulonglong2*>(GRt + (size_t)r * HD + kg * 4);
    rp0 = 0ull; rp1 = 0ull;
#pragma unroll
    for (int sg = 0; sg < CW; ++sg) {
        const bool bit = (w[sg] & lmask) != 0u;
        const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(bit ? prow1 : prow0);
        const u64 v0 = gate2(add2(p.x, Q[sg][0]), add2(g.x, GC[sg][0]));
        const u64 v1 = gate2(add2(p.y, Q[sg][1]), add2(g.y, GC[sg][1]));
        col[sg][0] = add2(col[sg][0], v0); col[sg][1] = add2(col[sg][1], v1);
        rp0 = add2(rp0, v0); rp1 = add2(rp1, v1);
        const float lf = bit ? 1.f : 0.f;
        const u64 l2 = pk2(lf, lf);
        lacc[0] = fma2(l2, v0, lacc[0]); lacc[1] = fma2(l2, v1, lacc[1]);
    }
}

template <int CW, bool ACCUM>
__device__ __forceinline__ void sweep2_bwd(const float* P01, const float* GRt, const uint32_t* bits, int WP, int nrows,
                                           int rg, int nrg, int kg, const u64 (&Q)[CW][2], const u64 (&GC)[CW][2],
                                           u64 (&col)[CW][2], u64 (&lacc)[2], float* rowacc, int lane) {
    const uint32_t lmask = 1u << lane;
    int r = rg;
    for (; r + 3 * nrg < nrows; r += 4 * nrg) {      // four rows per trip (see sweep2_fwd)
        u64 a0, a1, c0, c1, e0, e1, g0, g1;
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r, kg, lmask, Q, GC, col, lacc, a0, a1);
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r + nrg, kg, lmask, Q, GC, col, lacc, c0, c1);
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r + 2 * nrg, kg, lmask, Q, GC, col, lacc, e0, e1);
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r + 3 * nrg, kg, lmask, Q, GC, col, lacc, g0, g1);
        const float tot = reduce8(a0, a1, c0, c1, lane), tot2 = reduce8(e0, e1, g0, g1, lane);
        if ((lane & 3) == 0) {
            row_store<ACCUM>(rowacc, (lane & 16) ? r + nrg : r, kg * 4 + reduce8_channel(lane), tot);
            row_store<ACCUM>(rowacc, (lane & 16) ? r + 3 * nrg : r + 2 * nrg, kg * 4 + reduce8_channel(lane), tot2);
        }
    }
    for (; r + nrg < nrows; r += 2 * nrg) {
        u64 a0, a1, c0, c1;
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r, kg, lmask, Q, GC, col, lacc, a0, a1);
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r + nrg, kg, lmask, Q, GC, col, lacc, c0, c1);
        const float tot = reduce8(a0, a1, c0, c1, lane);
        if ((lane & 3) == 0) row_store<ACCUM>(rowacc, (lane & 16) ? r + nrg : r, kg * 4 + reduce8_channel(lane), tot);
    }
    if (r < nrows) {
        u64 a0, a1;
        sweep2_bwd_row<CW>(P01, GRt, bits, WP, r, kg, lmask, Q, GC, col, lacc, a0, a1);
        const float tot = reduce4(a0, a1, lane);
        if ((lane & 7) == 0) row_store<ACCUM>(rowacc, r, kg * 4 + reduce4_channel(lane), tot);
    }
}

// ---- column partials across the NRG row groups --------------------------------------------------
// Fixed-order pairwise tree through shared memory: (rg0 + rg2) + (rg1 + rg3) for NRG = 4, etc.
// `scratch` holds (NRG / 2 rounded up) * CW*32*20 floats.  Must be called by ALL threads of the
// CTA (it contains __syncthreads).  On return the warps with rg == 0 hold the totals in `col`.
template <int CW>
__device__ __forceinline__ void combine_cols(u64 (&col)[CW][2], float* scratch, int rg, int nrg, int kg, int lane) {
    const int slot_floats = CW * 32 * HD;
    for (int half = (nrg + 1) >> 1, n = nrg; n > 1; n = half, half = (half + 1) >> 1) {
        // row groups [half, n) store, row groups [0, n - half) add
        if (rg >= half && rg < n) {
            float* dst = scratch + (size_t)(rg - half) * slot_floats + kg * 4;
#pragma unroll
            for (int sg = 0; sg < CW; ++sg)
                *reinterpret_cast<ulonglong2*>(dst + (size_t)(sg * 32 + lane) * HD) = make_ulonglong2(col[sg][0], col[sg][1]);
        }
        __syncthreads();
        if (rg < n - half) {
            const float* src = scratch + (size_t)rg * slot_floats + kg * 4;
#pragma unroll
            for (int sg = 0; sg < CW; ++sg) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + (size_t)(sg * 32 + lane) * HD);
                col[sg][0] = add2(col[sg][0], v.x); col[sg][1] = add2(col[sg][1], v.y);
            }
        }
        __syncthreads();
    }
}

// The same tree with the column segments taken in two halves, for wide grids: `scratch` holds (NRG / 2 rounded up) *
// ceil(CW / 2)*32*20 floats.  Same additions in the same order as combine_cols.
template <int CW>
__device__ __forceinline__ void combine_cols_2pass(u64 (&col)[CW][2], float* scratch, int rg, int nrg, int kg, int lane) {
    constexpr int H = (CW + 1) / 2;
    const int slot_floats = H * 32 * HD;
#pragma unroll
    for (int s0 = 0; s0 < CW; s0 += H) {
        for (int half = (nrg + 1) >> 1, n = nrg; n > 1; n = half, half = (half + 1) >> 1) {
            if (rg >= half && rg < n) {
                float* dst = scratch + (size_t)(rg - half) * slot_floats + kg * 4;
#pragma unroll
                for (int sg = s0; sg < s0 + H && sg < CW; ++sg)
                    *reinterpret_cast<ulonglong2*>(dst + (size_t)((sg - s0) * 32 + lane) * HD) = make_ulonglong2(col[sg][0], col[sg][1]);
            }
            __syncthreads();
            if (rg < n - half) {
                const float* src = scratch + (size_t)rg * slot_floats + kg * 4;
#pragma unroll
                for (int sg = s0; sg < s0 + H && sg < CW; ++sg) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + (size_t)((sg - s0) * 32 + lane) * HD);
                    col[sg][0] = add2(col[sg][0], v.x); col[sg][1] = add2(col[sg][1], v.y);
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace hdgnn
