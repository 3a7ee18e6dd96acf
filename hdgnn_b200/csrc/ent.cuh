// Entity-grid kernels: the two big sweeps of a training step (model_2.py:161-188 forward,
// and its backward), one CTA per (row tile, commit).
//
//   ent_fwd:  RS_i = sum_{j != i} relu(pre_ij)   (complete rows of the tile)
//             CS_j = sum_{i in tile, i != j} relu(pre_ij)   (per-tile partial, (B,S,N,20))
//   ent_bwd:  v_ij = [pre_ij > 0] (GR_i + GC_j); the tile's contribution to the first-layer
//             weight gradients is reduced inside the CTA:
//                 gpart[b][s] = { db[k] = sum_i RSd_i[k],  dU[k] = sum_i x_i RSd_i[k],
//                                 dV[k] = sum_j x_j CSd_j[k],  LS[k] = sum_{l_ij = 1} v_ij[k] }
//             so nothing of size N x 20 leaves the SM.
//
// pre_ij[k] = x_i U[k] + x_j V[k] + b[k] + W_l[l_ij][k]  (rank-1 form of the 4 -> 20 layer,
// model_2.py:144,167-170; the entity-edge branch of model_4.py:219-225 uses U == V).
// The label tile (RT rows x pitch bytes, contiguous in HBM) is staged by one 1-D TMA bulk copy.
#pragma once
#include "sweep.cuh"

namespace hdgnn {

struct EntArgs {
    const uint8_t* lab;   // (B, N, pitch)
    int pitch, N, RT, S;
    const float* params;
    const float* x;       // (B, N)
    int o_u, o_v, o_b, o_l;
    const float* GR;      // bwd: (B,N,20) d/dRS
    const float* GC;      // bwd: (B,N,20) d/dCS
    float* RS;            // fwd: (B,N,20)
    float* CSp;           // fwd: (B,S,N,20)
    float* gpart;         // bwd: (B,S,80)
};

constexpr int CP_STRIDE = 21;                      // padded [lane][k] stride of a column partial
constexpr int CP_WARP = 32 * CP_STRIDE;

__host__ __device__ inline size_t ent_smem_bytes(int NW, int N, int RT, int pitch, bool bwd) {
    size_t off = round_up(RT * pitch, 128);
    off += (size_t)RT * HD * 4 * (bwd ? 5 : 4);            // P0, P1, rowacc, diag (, GR)
    off += (size_t)NW * CP_WARP * 4;                       // column partials
    if (bwd) off += (size_t)round_up(N, 32) * HD * 4 + (size_t)round_up(N, 32) * 4 + (size_t)(NW + 8 * 3) * HD * 4;
    return off + 16;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) ent_fwd_kernel(const EntArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, RT = a.RT, pitch = a.pitch;
    const int b = blockIdx.y, s = blockIdx.x, r0 = s * RT, nrows = min(RT, N - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* stage = smem;
    float* P01 = reinterpret_cast<float*>(smem + round_up(RT * pitch, 128));
    float* rowacc = P01 + 2 * RT * HD;
    float* diag = rowacc + RT * HD;
    float* cpart = diag + RT * HD;
    uint64_t* bar = reinterpret_cast<uint64_t*>(cpart + NW * CP_WARP);

    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)nrows * pitch;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(stage, a.lab + ((size_t)b * N + r0) * pitch, bytes, bar);
    }
    const float* par = a.params;
    const float* xb = a.x + (size_t)b * N;
    for (int idx = tid; idx < nrows * HD; idx += NW * 32) {
        const int r = idx / HD, k = idx - r * HD;
        const float xi = xb[r0 + r];
        const float p0 = fmaf(xi, par[a.o_u + k], par[a.o_b + k] + par[a.o_l + k]);
        P01[idx] = p0;
        P01[RT * HD + idx] = p0 + (par[a.o_l + HD + k] - par[a.o_l + k]);
        rowacc[idx] = 0.f;
        diag[idx] = fmaxf(p0 + xi * par[a.o_v + k], 0.f);
    }
    float V[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) V[k] = par[a.o_v + k];
    const int ridx = reduce20_index(lane);
    mbar_wait(bar, 0);
    __syncthreads();

    const int ncb = (N + 31) >> 5;
    float* csp = a.CSp + ((size_t)b * a.S + s) * N * HD;
    for (int cb = 0; cb < ncb; ++cb) {
        const int j = cb * 32 + lane;
        const bool ok = j < N;
        float Q[HD], col[HD];
        const float xj = ok ? xb[j] : 0.f;
#pragma unroll
        for (int k = 0; k < HD; ++k) { Q[k] = ok ? xj * V[k] : NEG_BIG; col[k] = 0.f; }
        sweep_fwd_block(P01, RT, stage + j, pitch, r0, j, ok, nrows, warp, NW, Q, col, rowacc, lane, ridx);
        float* dst = cpart + warp * CP_WARP + lane * CP_STRIDE;
#pragma unroll
        for (int k = 0; k < HD; ++k) dst[k] = col[k];
        __syncthreads();
        for (int idx = tid; idx < 32 * HD; idx += NW * 32) {
            const int ln = idx / HD, k = idx - ln * HD, jj = cb * 32 + ln;
            if (jj < N) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) v += cpart[w * CP_WARP + ln * CP_STRIDE + k];
                const int r = jj - r0;
                if (r >= 0 && r < nrows) v -= diag[r * HD + k];
                csp[(size_t)jj * HD + k] = v;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nrows * HD; idx += NW * 32)
        a.RS[((size_t)b * N + r0) * HD + idx] = rowacc[idx] - diag[idx];
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) ent_bwd_kernel(const EntArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, RT = a.RT, pitch = a.pitch;
    const int b = blockIdx.y, s = blockIdx.x, r0 = s * RT, nrows = min(RT, N - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = round_up(N, 32);
    uint8_t* stage = smem;
    float* P01 = reinterpret_cast<float*>(smem + round_up(RT * pitch, 128));
    float* rowacc = P01 + 2 * RT * HD;
    float* dgv = rowacc + RT * HD;
    float* GRt = dgv + RT * HD;
    float* cpart = GRt + RT * HD;
    float* CSd = cpart + NW * CP_WARP;          // [NP][20]
    float* xs = CSd + (size_t)NP * HD;           // [NP]
    float* lsw = xs + NP;                        // [NW][20]
    float* red = lsw + NW * HD;                  // [3][8][20]
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 3 * 8 * HD);

    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)nrows * pitch;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(stage, a.lab + ((size_t)b * N + r0) * pitch, bytes, bar);
    }
    const float* par = a.params;
    const float* xb = a.x + (size_t)b * N;
    const float* grb = a.GR + (size_t)b * N * HD;
    const float* gcb = a.GC + (size_t)b * N * HD;
    for (int idx = tid; idx < nrows * HD; idx += NW * 32) {
        const int r = idx / HD, k = idx - r * HD, i = r0 + r;
        const float xi = xb[i];
        const float p0 = fmaf(xi, par[a.o_u + k], par[a.o_b + k] + par[a.o_l + k]);
        P01[idx] = p0;
        P01[RT * HD + idx] = p0 + (par[a.o_l + HD + k] - par[a.o_l + k]);
        rowacc[idx] = 0.f;
        const float g = grb[(size_t)i * HD + k];
        GRt[idx] = g;
        dgv[idx] = (p0 + xi * par[a.o_v + k]) > 0.f ? g + gcb[(size_t)i * HD + k] : 0.f;
    }
    for (int j = tid; j < NP; j += NW * 32) xs[j] = j < N ? xb[j] : 0.f;
    float V[HD], lacc[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) { V[k] = par[a.o_v + k]; lacc[k] = 0.f; }
    const int ridx = reduce20_index(lane);
    mbar_wait(bar, 0);
    __syncthreads();

    const int ncb = (N + 31) >> 5;
    for (int cb = 0; cb < ncb; ++cb) {
        const int j = cb * 32 + lane;
        const bool ok = j < N;
        float Q[HD], GC[HD], col[HD];
        const float xj = ok ? xb[j] : 0.f;
        if (ok) load20(GC, gcb + (size_t)j * HD);
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            Q[k] = ok ? xj * V[k] : NEG_BIG;
            if (!ok) GC[k] = 0.f;
            col[k] = 0.f;
        }
        sweep_bwd_block(P01, GRt, RT, stage + j, pitch, r0, j, ok, nrows, warp, NW, Q, GC, col, lacc, rowacc, lane, ridx);
        float* dst = cpart + warp * CP_WARP + lane * CP_STRIDE;
#pragma unroll
        for (int k = 0; k < HD; ++k) dst[k] = col[k];
        __syncthreads();
        for (int idx = tid; idx < 32 * HD; idx += NW * 32) {
            const int ln = idx / HD, k = idx - ln * HD, jj = cb * 32 + ln;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += cpart[w * CP_WARP + ln * CP_STRIDE + k];
            const int r = jj - r0;
            if (r >= 0 && r < nrows) v -= dgv[r * HD + k];
            CSd[(size_t)jj * HD + k] = jj < N ? v : 0.f;
        }
        __syncthreads();
    }
    // label-1 sums: warp totals, then warps in order
    {
        const float t = warp_reduce20(lacc, lane);
        if (ridx >= 0) lsw[warp * HD + ridx] = t;
    }
    __syncthreads();
    // db, dU over the tile's rows; dV over all columns: 8 interleaved partials per channel, fixed order
    if (tid < 8 * HD) {
        const int p = tid / HD, k = tid - p * HD;
        float sb = 0.f, su = 0.f, sv = 0.f;
        for (int r = p; r < nrows; r += 8) {
            const float rs = rowacc[r * HD + k] - dgv[r * HD + k];
            sb += rs;
            su = fmaf(xs[r0 + r], rs, su);
        }
        for (int j = p; j < N; j += 8) sv = fmaf(xs[j], CSd[(size_t)j * HD + k], sv);
        red[(0 * 8 + p) * HD + k] = sb; red[(1 * 8 + p) * HD + k] = su; red[(2 * 8 + p) * HD + k] = sv;
    }
    __syncthreads();
    if (tid < HD) {
        const int k = tid;
        float sb = 0.f, su = 0.f, sv = 0.f, ls = 0.f;
        for (int p = 0; p < 8; ++p) { sb += red[(0 * 8 + p) * HD + k]; su += red[(1 * 8 + p) * HD + k]; sv += red[(2 * 8 + p) * HD + k]; }
        for (int w = 0; w < NW; ++w) ls += lsw[w * HD + k];
        float* gp = a.gpart + ((size_t)b * a.S + s) * 4 * HD;
        gp[k] = sb; gp[HD + k] = su; gp[2 * HD + k] = sv; gp[3 * HD + k] = ls;
    }
}

}  // namespace hdgnn
