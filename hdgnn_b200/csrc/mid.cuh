// mid: everything of a training step that is per-commit and smaller than the entity sweeps, fused
// into ONE kernel with one CTA per commit and every intermediate in shared memory:
//
//   entity-state MLP (model_2.py:181-205)  ->  entity->hunk pooling (model_2.py:146-150, index
//   semantics of utils2.py:111-137)  ->  hunk pair layer + "edge translation" (model_2.py:245-277)
//   ->  relation head + softmax + cross-entropy (model_2.py:304-324, 115-118)
//   ->  [training] the hand-written backward of all of the above down to d/d(entity effect sums).
//
// Only RS1 / CS1p (from ent_fwd) are read and GE (for ent_bwd), the per-commit gradient partials,
// probs / logits and the CE partial are written.  All reductions run in a fixed order.
#pragma once
#include "node.cuh"
#include "sweep.cuh"
#include "ent.cuh"

namespace hdgnn {

constexpr int MID_NW = 12;
constexpr int MID_THREADS = MID_NW * 32;
constexpr int MCH = 128;                 // entity nodes per chunk in the node-MLP phases

struct MidArgs {
    int Ne, Nc, Se, ent, train;
    const uint8_t* adj; int pe;
    const uint8_t* Y; int pc;
    const float* x; const int* hmap; const int* L;
    const float* params; ParamOff po;
    const float* RS1; const float* CS1p;     // (B,Ne,20), (B,Se,Ne,20)
    const float* soft; float* dsoft;         // (B,Ne,Ne,2) variant 4, else null
    float* logits; float* probs;             // (B,2,Ncr) or null
    float* cep;                              // (B) sum of CE over the commit's pairs
    float scale;                             // dL/dlogit scale: 10 / (B_global * Ncr)
    float* GE;                               // (B,Ne,20) d/dS1
    float* gpart; int total;                 // (B,total)
    float* dbg;                              // debug dumps (HDGNN_F_DEBUG) or null; layout below
};
// debug dump layout per commit (floats): S1[Ne*20] X2[Ne] NB[Nc*4] RS3[Nc*20] CS3[Nc*20] PR[Nc*20] PC[Nc*20]
// DNB[Nc*4] DX2[Ne]
__host__ __device__ inline size_t mid_dbg_floats(int Ne, int Nc) { return (size_t)Ne * 22 + (size_t)Nc * 88; }

struct MidSmem {
    // offsets in floats
    int W5, b5, U1, c1, u2, c2, V1, d1, W2, b2, G1, g1b, G2, gb2, gam, Dh, Dg;
    int x, x2, hm, SP, TP, dl, dx2, nb, dnb, cpart, red, uni, total;
};

__host__ __device__ inline MidSmem mid_layout(int Ne, int Nc) {
    MidSmem m;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    m.W5 = take(400); m.b5 = take(20); m.U1 = take(420); m.c1 = take(20); m.u2 = take(20); m.c2 = take(4);
    m.V1 = take(200); m.d1 = take(20); m.W2 = take(400); m.b2 = take(20); m.G1 = take(440); m.g1b = take(20);
    m.G2 = take(40); m.gb2 = take(4); m.gam = take(20); m.Dh = take(20); m.Dg = take(20);
    m.x = take(Ne); m.x2 = take(Ne); m.hm = take(Ne); m.SP = take(4 * Ne); m.TP = take(4 * Ne); m.dl = take(4 * Ne);
    m.dx2 = take(Ne); m.nb = take(4 * Nc); m.dnb = take(4 * Nc);
    m.cpart = take(MID_NW * CP_WARP); m.red = take(64 + MID_NW * HD + 64);
    m.uni = o;
    const int ent_phase = 5 * MCH * HD + MCH;
    const int hunk_phase = 12 * Nc * HD;
    o += ent_phase > hunk_phase ? ent_phase : hunk_phase;
    m.total = o;
    return m;
}
__host__ __device__ inline size_t mid_smem_bytes(int Ne, int Nc) { return (size_t)mid_layout(Ne, Nc).total * 4 + 16; }

// fixed-order block sum for MID_THREADS threads; every thread gets the result
__device__ __forceinline__ float mid_block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < MID_NW; ++w) t += scratch[w];
    return t;
}

// out[n][m] = bias_scale * bias[m] + sum_q in[n][q] * W[q][m]     (n < nn; 20 x 20, row-major W)
__device__ __forceinline__ void mm20(float* out, const float* in, const float* W, const float* bias, float bias_scale, int nn) {
    for (int idx = threadIdx.x; idx < nn * HD; idx += MID_THREADS) {
        const int n = idx / HD, m = idx - n * HD;
        float acc = bias ? bias_scale * bias[m] : 0.f;
#pragma unroll
        for (int q = 0; q < HD; ++q) acc = fmaf(in[n * HD + q], W[q * HD + m], acc);
        out[idx] = acc;
    }
}
// out[n][q] = sum_m W[q][m] * in[n][m]      (multiply by W^T)
__device__ __forceinline__ void mm20t(float* out, const float* in, const float* W, int nn) {
    for (int idx = threadIdx.x; idx < nn * HD; idx += MID_THREADS) {
        const int n = idx / HD, q = idx - n * HD;
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < HD; ++m) acc = fmaf(W[q * HD + m], in[n * HD + m], acc);
        out[idx] = acc;
    }
}

template <bool TRAIN, bool LOGITS>
__global__ void __launch_bounds__(MID_THREADS, 1) mid_kernel(const MidArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int Ne = a.Ne, Nc = a.Nc, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const MidSmem L_ = mid_layout(Ne, Nc);
    float* W5 = sm + L_.W5; float* b5 = sm + L_.b5; float* U1 = sm + L_.U1; float* c1 = sm + L_.c1;
    float* u2 = sm + L_.u2; float* c2 = sm + L_.c2; float* V1 = sm + L_.V1; float* d1 = sm + L_.d1;
    float* W2 = sm + L_.W2; float* b2 = sm + L_.b2; float* G1 = sm + L_.G1; float* g1b = sm + L_.g1b;
    float* G2 = sm + L_.G2; float* gb2 = sm + L_.gb2; float* gam = sm + L_.gam; float* Dh = sm + L_.Dh; float* Dg = sm + L_.Dg;
    float* xs = sm + L_.x; float* x2 = sm + L_.x2; int* hm = reinterpret_cast<int*>(sm + L_.hm);
    float* SP = sm + L_.SP; float* TP = sm + L_.TP; float* dl = sm + L_.dl; float* dx2 = sm + L_.dx2;
    float* nb = sm + L_.nb; float* dnb = sm + L_.dnb; float* cpart = sm + L_.cpart; float* red = sm + L_.red;
    float* uni = sm + L_.uni;
    const float* par = a.params;
    const ParamOff& po = a.po;
    const int ridx = reduce20_index(lane);
    float* dbg = a.dbg ? a.dbg + (size_t)b * mid_dbg_floats(Ne, Nc) : nullptr;
    float* gp = a.gpart ? a.gpart + (size_t)b * a.total : nullptr;

    // ---------------- A. weights and per-commit vectors -> shared memory ------------------------
    if (a.ent) {
        copy_to_smem(W5, par + po.ent_w5, 400); copy_to_smem(b5, par + po.ent_b5, 20);
        copy_to_smem(U1, par + po.nod_w1, 420); copy_to_smem(c1, par + po.nod_b1, 20);
        copy_to_smem(u2, par + po.nod_w2, 20);  copy_to_smem(c2, par + po.nod_b2, 1);
    }
    copy_to_smem(V1, par + po.hnk_w1, 200); copy_to_smem(d1, par + po.hnk_b1, 20);
    copy_to_smem(W2, par + po.hnk_w2, 400); copy_to_smem(b2, par + po.hnk_b2, 20);
    copy_to_smem(G1, par + po.scr_w1, 440); copy_to_smem(g1b, par + po.scr_b1, 20);
    copy_to_smem(G2, par + po.scr_w2, 40);  copy_to_smem(gb2, par + po.scr_b2, 2);
    if (tid < HD) {
        gam[tid] = par[po.scr_w2 + 2 * tid + 1] - par[po.scr_w2 + 2 * tid];
        Dh[tid] = par[po.hnk_w1 + 9 * HD + tid] - par[po.hnk_w1 + 8 * HD + tid];
        Dg[tid] = par[po.scr_w1 + HD + tid] - par[po.scr_w1 + tid];
    }
    const int Lb = a.L[b];
    for (int i = tid; i < Ne; i += MID_THREADS) {
        const int h = a.hmap[(size_t)b * Ne + i];
        hm[i] = (h >= 0 && h < Nc) ? h : -1;
        const float xv = a.x[(size_t)b * Ne + i];
        xs[i] = xv;
        if (!a.ent) x2[i] = xv;
    }
    __syncthreads();

    // ---------------- B/C. entity-state MLP forward, chunks of MCH nodes -------------------------
    if (a.ent) {
        float* sS = uni; float* sE = sS + MCH * HD; float* sZ = sE + MCH * HD;
        const float nb5 = 2.f * (float)(Ne - 1);
        for (int c0 = 0; c0 < Ne; c0 += MCH) {
            const int nn = min(MCH, Ne - c0);
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const size_t g = ((size_t)b * Ne + c0) * HD + idx;
                float v = a.RS1[g];
                for (int s = 0; s < a.Se; ++s) v += a.CS1p[((size_t)b * a.Se + s) * Ne * HD + (size_t)c0 * HD + idx];
                sS[idx] = v;
                if (dbg) dbg[(size_t)c0 * HD + idx] = v;
            }
            __syncthreads();
            mm20(sE, sS, W5, b5, nb5, nn);
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const int n = idx / HD, k = idx - n * HD;
                float acc = fmaf(xs[c0 + n], U1[k], c1[k]);
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(sE[n * HD + m], U1[(1 + m) * HD + k], acc);
                sZ[idx] = fmaxf(acc, 0.f);
            }
            __syncthreads();
            for (int n = tid; n < nn; n += MID_THREADS) {
                float acc = c2[0];
#pragma unroll
                for (int k = 0; k < HD; ++k) acc = fmaf(sZ[n * HD + k], u2[k], acc);
                x2[c0 + n] = fmaxf(acc, 0.f);
            }
            __syncthreads();
        }
    }
    if (dbg) for (int i = tid; i < Ne; i += MID_THREADS) dbg[(size_t)Ne * HD + i] = x2[i];

    // ---------------- D. pooling forward -------------------------------------------------------
    // B2[q] = [x2_gi, x2_gj, e0, e1] over the Ne-grid enumeration q; the L x L local grid selects
    // q = li (L-1) + lj - [lj > li]  (utils2.py:123-137, quirk Q3).  SP[li] = row sums, TP[lj] = column sums.
    {
        const int nm1 = Ne - 1;
        const float inv = 1.f / (float)nm1;
        const bool ident = Lb == Ne;
        for (int i = tid; i < 4 * Ne; i += MID_THREADS) SP[i] = 0.f;
        __syncthreads();
        const int nseg = (Lb + 31) >> 5;
        for (int seg = 0; seg < nseg; ++seg) {
            const int lj = seg * 32 + lane;
            float c0 = 0.f, c1_ = 0.f, c2_ = 0.f, c3 = 0.f;
            for (int li = warp; li < Lb; li += MID_NW) {
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (lj < Lb && lj != li) {
                    int gi = li, gj = lj;
                    if (!ident) unflat_pair(li * (Lb - 1) + lj - (lj > li), nm1, inv, gi, gj);
                    if (a.soft) {
                        const float2 s2 = reinterpret_cast<const float2*>(a.soft)[((size_t)b * Ne + gi) * Ne + gj];
                        v2 = s2.x; v3 = s2.y;
                    } else {
                        v3 = a.adj[((size_t)b * Ne + gi) * a.pe + gj] != 0 ? 1.f : 0.f;
                        v2 = 1.f - v3;
                    }
                    v0 = x2[gi]; v1 = x2[gj];
                }
                c0 += v0; c1_ += v1; c2_ += v2; c3 += v3;
                const float r0 = warp_sum(v0), r1 = warp_sum(v1), r2 = warp_sum(v2), r3 = warp_sum(v3);
                if (lane == 0) { SP[4 * li] += r0; SP[4 * li + 1] += r1; SP[4 * li + 2] += r2; SP[4 * li + 3] += r3; }
            }
            float* dst = cpart + warp * 128 + lane * 4;
            dst[0] = c0; dst[1] = c1_; dst[2] = c2_; dst[3] = c3;
            __syncthreads();
            if (tid < 128) {
                const int ln = tid >> 2, ch = tid & 3;
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < MID_NW; ++w) v += cpart[w * 128 + ln * 4 + ch];
                if (seg * 32 + ln < Ne) TP[4 * (seg * 32 + ln) + ch] = v;
            }
            __syncthreads();
        }
        // segmented reduce by hunk id in ascending entity-line order
        for (int idx = tid; idx < Nc * 4; idx += MID_THREADS) {
            const int c = idx >> 2, ch = idx & 3;
            float acc = 0.f;
            for (int i = 0; i < Lb; ++i)
                if (hm[i] == c) acc += SP[4 * i + ch] + TP[4 * i + ch];
            nb[idx] = acc;
            if (dbg) dbg[(size_t)Ne * 21 + idx] = acc;
        }
        __syncthreads();
    }

    // ---------------- hunk-stage tables (union region) ---------------------------------------------
    const int T = Nc * HD;
    float* PH01 = uni;              // [2][Nc][20]
    float* QH = uni + 2 * T;
    float* RS3 = uni + 3 * T;
    float* CS3 = uni + 4 * T;
    float* rr = uni + 5 * T;        // r, later GC
    float* cc = uni + 6 * T;
    float* PR01 = uni + 7 * T;      // [2][Nc][20]; PR1 half later holds GR
    float* PC = uni + 9 * T;        // later dc
    float* RSm = uni + 10 * T;      // later RS3d
    float* CSm = uni + 11 * T;      // later CS3d
    for (int idx = tid; idx < T; idx += MID_THREADS) {
        const int c = idx / HD, k = idx - c * HD;
        float p = d1[k] + V1[8 * HD + k], q = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            p = fmaf(nb[4 * c + ch], V1[ch * HD + k], p);
            q = fmaf(nb[4 * c + ch], V1[(4 + ch) * HD + k], q);
        }
        PH01[idx] = p; PH01[T + idx] = p + Dh[k]; QH[idx] = q;
        RS3[idx] = 0.f;
    }
    __syncthreads();

    // ---------------- E. hunk pair layer forward: row / column sums ---------------------------------
    const uint8_t* Yb = a.Y + (size_t)b * Nc * a.pc;
    const int ncb = (Nc + 31) >> 5;
    for (int cb = 0; cb < ncb; ++cb) {
        const int j = cb * 32 + lane;
        const bool ok = j < Nc;
        float Q[HD], col[HD];
        if (ok) load20(Q, QH + j * HD);
#pragma unroll
        for (int k = 0; k < HD; ++k) { if (!ok) Q[k] = NEG_BIG; col[k] = 0.f; }
        sweep_fwd_block(PH01, Nc, Yb + j, a.pc, 0, j, ok, Nc, warp, MID_NW, Q, col, RS3, lane, ridx);
        float* dst = cpart + warp * CP_WARP + lane * CP_STRIDE;
#pragma unroll
        for (int k = 0; k < HD; ++k) dst[k] = col[k];
        __syncthreads();
        for (int idx = tid; idx < 32 * HD; idx += MID_THREADS) {
            const int ln = idx / HD, k = idx - ln * HD, jj = cb * 32 + ln;
            if (jj < Nc) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < MID_NW; ++w) v += cpart[w * CP_WARP + ln * CP_STRIDE + k];
                CS3[jj * HD + k] = v;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < T; idx += MID_THREADS) {       // remove the diagonal pair (l = 0)
        const float d = fmaxf(PH01[idx] + QH[idx], 0.f);
        RS3[idx] -= d; CS3[idx] -= d;
        if (dbg) { dbg[(size_t)Ne * 21 + Nc * 4 + idx] = RS3[idx]; dbg[(size_t)Ne * 21 + Nc * 4 + T + idx] = CS3[idx]; }
    }
    __syncthreads();

    // ---------------- F. linear second layer on the sums + head tables (model_2.py:263-275, 311-315) --
    mm20(rr, RS3, W2, b2, (float)(Nc - 1), Nc);
    mm20(cc, CS3, W2, b2, (float)(Nc - 1), Nc);
    __syncthreads();
    for (int idx = tid; idx < T; idx += MID_THREADS) {
        const int n = idx / HD, k = idx - n * HD;
        float p = g1b[k] + G1[k], q = 0.f;
#pragma unroll
        for (int m = 0; m < HD; ++m) {
            p = fmaf(rr[n * HD + m], G1[(2 + m) * HD + k], p);
            q = fmaf(cc[n * HD + m], G1[(2 + m) * HD + k], q);
        }
        PR01[idx] = p; PR01[T + idx] = p + Dg[k]; PC[idx] = q;
        RSm[idx] = 0.f;
        if (dbg) { dbg[(size_t)Ne * 21 + Nc * 4 + 2 * T + idx] = p; dbg[(size_t)Ne * 21 + Nc * 4 + 3 * T + idx] = q; }
    }
    __syncthreads();

    // ---------------- G. relation head: logits, softmax, CE, and the delta sums ------------------------
    float ce_acc = 0.f, d_acc = 0.f;
    float lsm[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) lsm[k] = 0.f;
    {
        const size_t npair = (size_t)Nc * (Nc - 1);
        const float bd = gb2[1] - gb2[0], b20 = gb2[0];
        for (int cb = 0; cb < ncb; ++cb) {
            const int j = cb * 32 + lane;
            const bool ok = j < Nc;
            float Q[HD], col[HD];
            if (ok) load20(Q, PC + j * HD);
#pragma unroll
            for (int k = 0; k < HD; ++k) { if (!ok) Q[k] = NEG_BIG; col[k] = 0.f; }
            for (int r = warp; r < Nc; r += MID_NW) {
                const bool valid = ok && j != r;
                const bool lab = valid && Yb[(size_t)r * a.pc + j] != 0;
                float P[HD];
                load20(P, PR01 + ((lab ? Nc : 0) + r) * HD);
                float t[HD];
                float d = bd, l0 = b20;
#pragma unroll
                for (int k = 0; k < HD; ++k) {
                    t[k] = P[k] + Q[k];
                    const float h = fmaxf(t[k], 0.f);
                    d = fmaf(h, gam[k], d);
                    if (LOGITS) l0 = fmaf(h, G2[2 * k], l0);
                }
                const float e = expf(-fabsf(d));
                const float inv = 1.f / (1.f + e);
                const float p1 = d >= 0.f ? inv : e * inv, p0 = d >= 0.f ? e * inv : inv;
                if (valid) {
                    const size_t q = (size_t)r * (Nc - 1) + j - (j > r);
                    if (a.probs) {
                        a.probs[((size_t)b * 2 + 0) * npair + q] = p0;
                        a.probs[((size_t)b * 2 + 1) * npair + q] = p1;
                    }
                    if (LOGITS) {
                        a.logits[((size_t)b * 2 + 0) * npair + q] = l0;
                        a.logits[((size_t)b * 2 + 1) * npair + q] = l0 + d;
                    }
                    const float z = lab ? -d : d;
                    ce_acc += fmaxf(z, 0.f) + log1pf(e);
                }
                if (TRAIN) {
                    const float dlt = valid ? a.scale * (p1 - (lab ? 1.f : 0.f)) : 0.f;
                    d_acc += dlt;
                    float v[HD];
#pragma unroll
                    for (int k = 0; k < HD; ++k) {
                        v[k] = t[k] > 0.f ? dlt : 0.f;
                        col[k] += v[k];
                        if (lab) lsm[k] += v[k];
                    }
                    const float tot = warp_reduce20(v, lane);
                    if (ridx >= 0) RSm[r * HD + ridx] += tot;
                }
            }
            if (TRAIN) {
                float* dst = cpart + warp * CP_WARP + lane * CP_STRIDE;
#pragma unroll
                for (int k = 0; k < HD; ++k) dst[k] = col[k];
                __syncthreads();
                for (int idx = tid; idx < 32 * HD; idx += MID_THREADS) {
                    const int ln = idx / HD, k = idx - ln * HD, jj = cb * 32 + ln;
                    if (jj < Nc) {
                        float v = 0.f;
#pragma unroll
                        for (int w = 0; w < MID_NW; ++w) v += cpart[w * CP_WARP + ln * CP_STRIDE + k];
                        CSm[jj * HD + k] = v;
                    }
                }
                __syncthreads();
            }
        }
    }
    {
        const float ce_tot = mid_block_sum(ce_acc, red);
        if (tid == 0 && a.cep) a.cep[b] = ce_tot;
    }
    if (!TRAIN) return;

    // ---------------- H. head backward (node level) ------------------------------------------------------
    float* lsw = red + 64;          // [MID_NW][20]
    float* misc = red + 64 + MID_NW * HD;   // [0..19] LSm then LS4, [40] dsum
    {
        const float d_tot = mid_block_sum(d_acc, red);
        const float t = warp_reduce20(lsm, lane);
        if (ridx >= 0) lsw[warp * HD + ridx] = t;
        __syncthreads();
        if (tid < HD) {
            float ls = 0.f;
            for (int w = 0; w < MID_NW; ++w) ls += lsw[w * HD + tid];
            misc[tid] = ls;
        }
        if (tid == 0) misc[40] = d_tot;
        __syncthreads();
        // HS[k] = sum_pairs relu(pre)[k] * delta  via  relu(pre) = m * (PR0_i + l Dg + PC_j)
        if (tid < HD) {
            const int k = tid;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) {
                acc = fmaf(PR01[n * HD + k], RSm[n * HD + k], acc);
                acc = fmaf(PC[n * HD + k], CSm[n * HD + k], acc);
            }
            acc = fmaf(Dg[k], misc[k], acc);
            gp[po.scr_w2 + 2 * k + 1] = acc;
            gp[po.scr_w2 + 2 * k] = -acc;
        }
        __syncthreads();
        for (int idx = tid; idx < T; idx += MID_THREADS) {       // RS4 = gam * RSm, CS4 = gam * CSm (in place)
            const int k = idx % HD;
            RSm[idx] *= gam[k]; CSm[idx] *= gam[k];
        }
        if (tid < HD) misc[tid] *= gam[tid];                     // LS4
        __syncthreads();
    }
    float* RS4 = RSm; float* CS4 = CSm;
    // scr_w1 rows 2.. : dG1e[m][k] = sum_n r[n][m] RS4[n][k] + c[n][m] CS4[n][k];  400 outputs
    for (int e = tid; e < 400 + HD; e += MID_THREADS) {
        if (e < 400) {
            const int m = e / HD, k = e - m * HD;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) {
                acc = fmaf(rr[n * HD + m], RS4[n * HD + k], acc);
                acc = fmaf(cc[n * HD + m], CS4[n * HD + k], acc);
            }
            gp[po.scr_w1 + 2 * HD + e] = acc;
        } else {                          // scr_b1 and the two label rows
            const int k = e - 400;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) acc += RS4[n * HD + k];
            gp[po.scr_b1 + k] = acc;
            gp[po.scr_w1 + HD + k] = misc[k];
            gp[po.scr_w1 + k] = acc - misc[k];
        }
    }
    if (tid == 0) { gp[po.scr_b2 + 1] = misc[40]; gp[po.scr_b2] = -misc[40]; }
    __syncthreads();
    // dr = RS4 G1e^T -> PR0 buffer ; dc = CS4 G1e^T -> PC buffer
    float* dr = PR01; float* dc = PC;
    mm20t(dr, RS4, G1 + 2 * HD, Nc);
    mm20t(dc, CS4, G1 + 2 * HD, Nc);
    __syncthreads();
    // hnk_w2[q][m] = sum_n RS3[n][q] dr[n][m] + CS3[n][q] dc[n][m] ; hnk_b2[m] = (Nc-1) sum_n (dr+dc)[n][m]
    for (int e = tid; e < 400 + HD; e += MID_THREADS) {
        if (e < 400) {
            const int q = e / HD, m = e - q * HD;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) {
                acc = fmaf(RS3[n * HD + q], dr[n * HD + m], acc);
                acc = fmaf(CS3[n * HD + q], dc[n * HD + m], acc);
            }
            gp[po.hnk_w2 + e] = acc;
        } else {
            const int m = e - 400;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) acc += dr[n * HD + m] + dc[n * HD + m];
            gp[po.hnk_b2 + m] = (float)(Nc - 1) * acc;
        }
    }
    float* GR = PR01 + T; float* GC = rr;
    __syncthreads();                 // rr (r) fully consumed above before it is overwritten by GC
    mm20t(GR, dr, W2, Nc);
    mm20t(GC, dc, W2, Nc);
    float* RS3d = RSm; float* CS3d = CSm;
    __syncthreads();
    for (int idx = tid; idx < T; idx += MID_THREADS) RS3d[idx] = 0.f;     // RS4 consumed (dr) -> reuse
    __syncthreads();

    // ---------------- I. hunk pair layer backward sweep ------------------------------------------------------
    float ls3[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) ls3[k] = 0.f;
    for (int cb = 0; cb < ncb; ++cb) {
        const int j = cb * 32 + lane;
        const bool ok = j < Nc;
        float Q[HD], GCr[HD], col[HD];
        if (ok) { load20(Q, QH + j * HD); load20(GCr, GC + j * HD); }
#pragma unroll
        for (int k = 0; k < HD; ++k) { if (!ok) { Q[k] = NEG_BIG; GCr[k] = 0.f; } col[k] = 0.f; }
        sweep_bwd_block(PH01, GR, Nc, Yb + j, a.pc, 0, j, ok, Nc, warp, MID_NW, Q, GCr, col, ls3, RS3d, lane, ridx);
        float* dst = cpart + warp * CP_WARP + lane * CP_STRIDE;
#pragma unroll
        for (int k = 0; k < HD; ++k) dst[k] = col[k];
        __syncthreads();
        for (int idx = tid; idx < 32 * HD; idx += MID_THREADS) {
            const int ln = idx / HD, k = idx - ln * HD, jj = cb * 32 + ln;
            if (jj < Nc) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < MID_NW; ++w) v += cpart[w * CP_WARP + ln * CP_STRIDE + k];
                CS3d[jj * HD + k] = v;
            }
        }
        __syncthreads();
    }
    {
        const float t = warp_reduce20(ls3, lane);
        if (ridx >= 0) lsw[warp * HD + ridx] = t;
    }
    for (int idx = tid; idx < T; idx += MID_THREADS) {       // diagonal pair: l = 0
        const float d = (PH01[idx] + QH[idx]) > 0.f ? GR[idx] + GC[idx] : 0.f;
        RS3d[idx] -= d; CS3d[idx] -= d;
    }
    __syncthreads();

    // ---------------- J. hunk first-layer weights, d/dnb -----------------------------------------------------
    for (int e = tid; e < 8 * HD + HD; e += MID_THREADS) {
        if (e < 8 * HD) {
            const int row = e / HD, k = e - row * HD, ch = row & 3;
            const float* src = row < 4 ? RS3d : CS3d;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) acc = fmaf(nb[4 * n + ch], src[n * HD + k], acc);
            gp[po.hnk_w1 + e] = acc;
        } else {
            const int k = e - 8 * HD;
            float acc = 0.f, ls = 0.f;
            for (int n = 0; n < Nc; ++n) acc += RS3d[n * HD + k];
            for (int w = 0; w < MID_NW; ++w) ls += lsw[w * HD + k];
            gp[po.hnk_b1 + k] = acc;
            gp[po.hnk_w1 + 9 * HD + k] = ls;
            gp[po.hnk_w1 + 8 * HD + k] = acc - ls;
        }
    }
    for (int idx = tid; idx < Nc * 4; idx += MID_THREADS) {
        const int n = idx >> 2, ch = idx & 3;
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            v = fmaf(V1[ch * HD + k], RS3d[n * HD + k], v);
            v = fmaf(V1[(4 + ch) * HD + k], CS3d[n * HD + k], v);
        }
        dnb[idx] = v;
        if (dbg) dbg[(size_t)Ne * 21 + Nc * 4 + 4 * T + idx] = v;
    }
    __syncthreads();
    if (!a.ent && !a.dsoft) return;

    // ---------------- K. pooling backward ---------------------------------------------------------------------
    // dB2[q] = dnb[hunk(li)] + dnb[hunk(lj)] for q < L(L-1);  dx2[gi] += dB2[q][0], dx2[gj] += dB2[q][1]
    for (int idx = tid; idx < Ne * 4; idx += MID_THREADS) {
        const int i = idx >> 2, ch = idx & 3;
        dl[idx] = (i < Lb && hm[i] >= 0) ? dnb[4 * hm[i] + ch] : 0.f;
    }
    for (int i = tid; i < Ne; i += MID_THREADS) dx2[i] = 0.f;
    __syncthreads();
    {
        const int lm1 = Lb - 1, nm1 = Ne - 1, qmax = Lb * lm1;
        const float invl = 1.f / (float)lm1;
        const bool ident = Lb == Ne;
        const int nseg = (Ne + 31) >> 5;
        float* TPc = TP;     // column part of dx2, [Ne]
        for (int seg = 0; seg < nseg; ++seg) {
            const int gj = seg * 32 + lane;
            float cacc = 0.f;
            for (int gi = warp; gi < Ne; gi += MID_NW) {
                float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                const bool inb = gj < Ne && gj != gi;
                if (inb) {
                    const int q = gi * nm1 + gj - (gj > gi);
                    if (q < qmax) {
                        int li = gi, lj = gj;
                        if (!ident) unflat_pair(q, lm1, invl, li, lj);
                        w0 = dl[4 * li] + dl[4 * lj];
                        w1 = dl[4 * li + 1] + dl[4 * lj + 1];
                        if (a.dsoft) { w2 = dl[4 * li + 2] + dl[4 * lj + 2]; w3 = dl[4 * li + 3] + dl[4 * lj + 3]; }
                    }
                    if (a.dsoft) reinterpret_cast<float2*>(a.dsoft)[((size_t)b * Ne + gi) * Ne + gj] = make_float2(w2, w3);
                }
                cacc += w1;
                const float r = warp_sum(w0);
                if (lane == 0) dx2[gi] += r;
            }
            cpart[warp * 32 + lane] = cacc;
            __syncthreads();
            if (tid < 32) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < MID_NW; ++w) v += cpart[w * 32 + tid];
                if (seg * 32 + tid < Ne) TPc[seg * 32 + tid] = v;
            }
            __syncthreads();
        }
        for (int n = tid; n < Ne; n += MID_THREADS) {
            dx2[n] += TPc[n];
            if (dbg) dbg[(size_t)Ne * 21 + Nc * 8 + 4 * T + n] = dx2[n];
        }
        __syncthreads();
    }
    if (!a.ent) return;

    // ---------------- L. entity-state MLP backward (model_2.py:190-205) and W5/b5 (model_2.py:172-175) ------------
    {
        float* sS = uni; float* sE = sS + MCH * HD; float* sZp = sE + MCH * HD; float* sDz = sZp + MCH * HD;
        float* sDE = sDz + MCH * HD; float* sdu = sDE + MCH * HD;
        const float nb5 = 2.f * (float)(Ne - 1);
        float acc3[3] = {0.f, 0.f, 0.f};      // entries e = tid + rep * MID_THREADS, e < 881
        for (int c0 = 0; c0 < Ne; c0 += MCH) {
            const int nn = min(MCH, Ne - c0);
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const size_t g = ((size_t)b * Ne + c0) * HD + idx;
                float v = a.RS1[g];
                for (int s = 0; s < a.Se; ++s) v += a.CS1p[((size_t)b * a.Se + s) * Ne * HD + (size_t)c0 * HD + idx];
                sS[idx] = v;
            }
            __syncthreads();
            mm20(sE, sS, W5, b5, nb5, nn);
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const int n = idx / HD, k = idx - n * HD;
                float acc = fmaf(xs[c0 + n], U1[k], c1[k]);
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(sE[n * HD + m], U1[(1 + m) * HD + k], acc);
                sZp[idx] = acc;
            }
            for (int n = tid; n < nn; n += MID_THREADS) sdu[n] = x2[c0 + n] > 0.f ? dx2[c0 + n] : 0.f;
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const int n = idx / HD, k = idx - n * HD;
                sDz[idx] = sZp[idx] > 0.f ? sdu[n] * u2[k] : 0.f;
            }
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {
                const int n = idx / HD, m = idx - n * HD;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < HD; ++k) acc = fmaf(U1[(1 + m) * HD + k], sDz[n * HD + k], acc);
                sDE[idx] = acc;
            }
            __syncthreads();
#pragma unroll
            for (int rep = 0; rep < 3; ++rep) {
                const int e = tid + rep * MID_THREADS;
                float acc = 0.f;
                if (e < 400) {                       // dU1[1+m][k]
                    const int m = e / HD, k = e - m * HD;
                    for (int n = 0; n < nn; ++n) acc = fmaf(sE[n * HD + m], sDz[n * HD + k], acc);
                } else if (e < 420) {                // dU1[0][k]
                    const int k = e - 400;
                    for (int n = 0; n < nn; ++n) acc = fmaf(xs[c0 + n], sDz[n * HD + k], acc);
                } else if (e < 440) {                // dc1[k]
                    const int k = e - 420;
                    for (int n = 0; n < nn; ++n) acc += sDz[n * HD + k];
                } else if (e < 460) {                // du2[k]
                    const int k = e - 440;
                    for (int n = 0; n < nn; ++n) acc = fmaf(sdu[n], fmaxf(sZp[n * HD + k], 0.f), acc);
                } else if (e == 460) {               // dc2
                    for (int n = 0; n < nn; ++n) acc += sdu[n];
                } else if (e < 861) {                // dW5[q][m]
                    const int q = (e - 461) / HD, m = (e - 461) - q * HD;
                    for (int n = 0; n < nn; ++n) acc = fmaf(sS[n * HD + q], sDE[n * HD + m], acc);
                } else if (e < 881) {                // db5[m]
                    const int m = e - 861;
                    for (int n = 0; n < nn; ++n) acc += sDE[n * HD + m];
                    acc *= nb5;
                }
                acc3[rep] += acc;
            }
            for (int idx = tid; idx < nn * HD; idx += MID_THREADS) {      // GE = dEbar W5^T
                const int n = idx / HD, q = idx - n * HD;
                float acc = 0.f;
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(W5[q * HD + m], sDE[n * HD + m], acc);
                a.GE[((size_t)b * Ne + c0) * HD + idx] = acc;
            }
            __syncthreads();
        }
#pragma unroll
        for (int rep = 0; rep < 3; ++rep) {
            const int e = tid + rep * MID_THREADS;
            const float v = acc3[rep];
            if (e < 400) gp[po.nod_w1 + HD + e] = v;
            else if (e < 420) gp[po.nod_w1 + (e - 400)] = v;
            else if (e < 440) gp[po.nod_b1 + (e - 420)] = v;
            else if (e < 460) gp[po.nod_w2 + (e - 440)] = v;
            else if (e == 460) gp[po.nod_b2] = v;
            else if (e < 861) gp[po.ent_w5 + (e - 461)] = v;
            else if (e < 881) gp[po.ent_b5 + (e - 861)] = v;
        }
    }
}

}  // namespace hdgnn
