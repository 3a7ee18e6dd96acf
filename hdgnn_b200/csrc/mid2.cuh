// mid2: everything of a training step that is per-commit and smaller than the entity sweeps, fused
// into ONE kernel with one CTA per commit and every intermediate in shared memory:
//
//   entity-state MLP (model_2.py:181-205)  ->  entity->hunk pooling (model_2.py:146-150, index
//   semantics of utils2.py:111-137)  ->  hunk pair layer + "edge translation" (model_2.py:245-277)
//   ->  relation head + softmax + cross-entropy (model_2.py:304-324, 115-118)
//   ->  [training] the hand-written backward of all of the above down to d/d(entity effect sums).
//
// Only RS1 / CS1p (from ent_fwd2) and the two label bitmaps are read (bitmaps: one TMA bulk copy
// each); GE (for ent_bwd2), the per-commit gradient partials, probs / logits and the CE partial are
// written.  All reductions run in a fixed order.
//
// Pooling: with all L = Ne index lines present the L x L local grid IS the Ne x Ne grid and the
// per-entity sums have the closed form {(Ne-1) x_i, X - x_i, (Ne-1) - deg_i, deg_i} (row / column
// degrees from popcounts of the bitmap); with L < Ne (quirk Q3: the local counter of
// utils2.py:123-137 runs with stride L-1) the L x L grid is gathered pair by pair from shared memory.
#pragma once
#include "sweep2.cuh"
#include "ent2.cuh"

namespace hdgnn {

constexpr int M2_NRG = 4;
constexpr int M2_NW = KG * M2_NRG;        // 20 warps
constexpr int M2_T = M2_NW * 32;          // 640 threads
constexpr int M2_CH = 128;                // entity nodes per chunk in the node-MLP phases

struct Mid2Args {
    int Ne, Nc, ent, R, SL;                  // R, SL: row-chunk decomposition of ent_fwd2 (ent2.cuh)
    const uint32_t* ebits; int WPe;          // (B,Ne,WPe)
    const uint32_t* ybits; int WPc;          // (B,Nc,WPc)
    const float* x; const int* hmap; const int* L;
    const float* params; ParamOff po;
    const float* RS1; const float* CS1p;     // (B,Ne,20), (B,SL,Ne,20)
    float* logits; float* probs;             // (B,2,Ncr) or null
    float* cep;                              // (B) sum of CE over the commit's pairs
    float scale;                             // dL/dlogit scale: 10 / (B_global * Ncr)
    float* GE;                               // (B,Ne,20) d/dS1
    float* gpart; int total;                 // (B,total)
    float* dbg;                              // debug dumps (HDGNN_F_DEBUG) or null; layout below
    long long* clk;                          // per-phase clock64 stamps (B,16) or null
};
// debug dump layout per commit (floats): S1[Ne*20] X2[Ne] NB[Nc*4] RS3[Nc*20] CS3[Nc*20] PR[Nc*20] PC[Nc*20]
// DNB[Nc*4] DX2[Ne]
__host__ __device__ inline size_t mid2_dbg_floats(int Ne, int Nc) { return (size_t)Ne * 22 + (size_t)Nc * 88; }

struct Mid2Smem {
    // offsets in floats
    int W5, b5, U1, c1, u2, c2, V1, d1, W2, b2, G1, g1b, G2, gb2, gam, Dh, Dg;
    int x, x2, hm, SP, TP, dl, dx2, nb, dnb, ebits, ybits, scratch, red, uni, total;
};

__host__ __device__ inline Mid2Smem mid2_layout(int Ne, int Nc, bool train) {
    Mid2Smem m;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 7) & ~7; return r; };      // 32-byte granules
    m.W5 = take(400); m.b5 = take(20); m.U1 = take(420); m.c1 = take(20); m.u2 = take(20); m.c2 = take(4);
    m.V1 = take(200); m.d1 = take(20); m.W2 = take(400); m.b2 = take(20); m.G1 = take(440); m.g1b = take(20);
    m.G2 = take(40); m.gb2 = take(4); m.gam = take(20); m.Dh = take(20); m.Dg = take(20);
    m.x = take(Ne); m.x2 = take(Ne); m.hm = take(Ne); m.SP = take(4 * Ne); m.TP = take(4 * Ne); m.dl = take(4 * Ne);
    m.dx2 = take(Ne); m.nb = take(4 * Nc); m.dnb = take(4 * Nc);
    m.ebits = take(Ne * bit_words(Ne)); m.ybits = take(Nc * bit_words(Nc));
    const int cwc = (Nc + 31) / 32;
    const int comb = (M2_NRG / 2) * cwc * 32 * HD, pool = 2 * 4 * 4 * Ne;     // column combine | pooling partials
    m.scratch = take(comb > pool ? comb : pool);
    m.red = take(64 + M2_NW * HD + 64);
    m.uni = o;
    const int ent_phase = 5 * M2_CH * HD + M2_CH;
    const int hunk_phase = 12 * Nc * HD + (train ? Nc * cwc * 32 : 0);
    o += ent_phase > hunk_phase ? ent_phase : hunk_phase;
    m.total = o;
    return m;
}
__host__ __device__ inline size_t mid2_smem_bytes(int Ne, int Nc, bool train) {
    return (size_t)mid2_layout(Ne, Nc, train).total * 4 + 16;
}

// fixed-order block sum for M2_T threads; every thread gets the result
__device__ __forceinline__ float mid2_block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < M2_NW; ++w) t += scratch[w];
    return t;
}

// out[n][m] = bias_scale * bias[m] + sum_q in[n][q] * W[q][m]     (n < nn; 20 x 20, row-major W)
__device__ __forceinline__ void m2_mm20(float* out, const float* in, const float* W, const float* bias, float bias_scale, int nn) {
    for (int idx = threadIdx.x; idx < nn * HD; idx += M2_T) {
        const int n = idx / HD, m = idx - n * HD;
        float acc = bias ? bias_scale * bias[m] : 0.f;
#pragma unroll
        for (int q = 0; q < HD; ++q) acc = fmaf(in[n * HD + q], W[q * HD + m], acc);
        out[idx] = acc;
    }
}
// out[n][q] = sum_m W[q][m] * in[n][m]      (multiply by W^T)
__device__ __forceinline__ void m2_mm20t(float* out, const float* in, const float* W, int nn) {
    for (int idx = threadIdx.x; idx < nn * HD; idx += M2_T) {
        const int n = idx / HD, q = idx - n * HD;
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < HD; ++m) acc = fmaf(W[q * HD + m], in[n * HD + m], acc);
        out[idx] = acc;
    }
}
// element k of row n of a [n][KG][2][4] table (half 0)
__device__ __forceinline__ int p01_idx(int n, int k) { return n * PROW + (k >> 2) * 8 + (k & 3); }

#define M2_PHASE(i) do { if (a.clk && tid == 0) a.clk[(size_t)b * 16 + (i)] = clock64(); } while (0)

template <int CWT, bool TRAIN>
__global__ void __launch_bounds__(M2_T, 1) mid2_kernel(const Mid2Args a) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    float* sm = reinterpret_cast<float*>(sm_raw);
    const int Ne = a.Ne, Nc = a.Nc, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kg = warp % KG, rg = warp / KG, k0 = kg * 4;
    const int WPe = a.WPe, WPc = a.WPc;
    const Mid2Smem L_ = mid2_layout(Ne, Nc, TRAIN);
    float* W5 = sm + L_.W5; float* b5 = sm + L_.b5; float* U1 = sm + L_.U1; float* c1 = sm + L_.c1;
    float* u2 = sm + L_.u2; float* c2 = sm + L_.c2; float* V1 = sm + L_.V1; float* d1 = sm + L_.d1;
    float* W2 = sm + L_.W2; float* b2 = sm + L_.b2; float* G1 = sm + L_.G1; float* g1b = sm + L_.g1b;
    float* G2 = sm + L_.G2; float* gb2 = sm + L_.gb2; float* gam = sm + L_.gam; float* Dh = sm + L_.Dh; float* Dg = sm + L_.Dg;
    float* xs = sm + L_.x; float* x2 = sm + L_.x2; int* hm = reinterpret_cast<int*>(sm + L_.hm);
    float* SP = sm + L_.SP; float* TP = sm + L_.TP; float* dl = sm + L_.dl; float* dx2 = sm + L_.dx2;
    float* nb = sm + L_.nb; float* dnb = sm + L_.dnb;
    uint32_t* ebits = reinterpret_cast<uint32_t*>(sm + L_.ebits);
    uint32_t* ybits = reinterpret_cast<uint32_t*>(sm + L_.ybits);
    float* scratch = sm + L_.scratch; float* red = sm + L_.red;
    float* uni = sm + L_.uni;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L_.total);
    const float* par = a.params;
    const ParamOff& po = a.po;
    const int ch = reduce4_channel(lane);
    const bool LOGITS = a.logits != nullptr;
    float* dbg = a.dbg ? a.dbg + (size_t)b * mid2_dbg_floats(Ne, Nc) : nullptr;
    float* gp = a.gpart ? a.gpart + (size_t)b * a.total : nullptr;
    M2_PHASE(0);

    // ---------------- A. label bitmaps (TMA), weights and per-commit vectors -> shared memory ------
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        const uint32_t be = (uint32_t)Ne * WPe * 4, bc = (uint32_t)Nc * WPc * 4;
        mbar_arrive_expect_tx(bar, be + bc);
        bulk_g2s(ebits, a.ebits + (size_t)b * Ne * WPe, be, bar);
        bulk_g2s(ybits, a.ybits + (size_t)b * Nc * WPc, bc, bar);
    }
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
    for (int i = tid; i < Ne; i += M2_T) {
        const int h = a.hmap[(size_t)b * Ne + i];
        hm[i] = (h >= 0 && h < Nc) ? h : -1;
        const float xv = a.x[(size_t)b * Ne + i];
        xs[i] = xv;
        if (!a.ent) x2[i] = xv;
    }
    __syncthreads();
    M2_PHASE(1);

    // ---------------- B/C. entity-state MLP forward, chunks of M2_CH nodes -------------------------
    const int nsl = a.ent ? ent2_slots(b, Ne, a.R) : 0;
    if (a.ent) {
        float* sS = uni; float* sE = sS + M2_CH * HD; float* sZ = sE + M2_CH * HD;
        const float nb5 = 2.f * (float)(Ne - 1);
        for (int c0 = 0; c0 < Ne; c0 += M2_CH) {
            const int nn = min(M2_CH, Ne - c0);
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const size_t g = ((size_t)b * Ne + c0) * HD + idx;
                float v = a.RS1[g];
                for (int s = 0; s < nsl; ++s) v += a.CS1p[((size_t)b * a.SL + s) * Ne * HD + (size_t)c0 * HD + idx];
                sS[idx] = v;
                if (dbg) dbg[(size_t)c0 * HD + idx] = v;
            }
            __syncthreads();
            m2_mm20(sE, sS, W5, b5, nb5, nn);
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const int n = idx / HD, k = idx - n * HD;
                float acc = fmaf(xs[c0 + n], U1[k], c1[k]);
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(sE[n * HD + m], U1[(1 + m) * HD + k], acc);
                sZ[idx] = fmaxf(acc, 0.f);
            }
            __syncthreads();
            for (int n = tid; n < nn; n += M2_T) {
                float acc = c2[0];
#pragma unroll
                for (int k = 0; k < HD; ++k) acc = fmaf(sZ[n * HD + k], u2[k], acc);
                x2[c0 + n] = fmaxf(acc, 0.f);
            }
            __syncthreads();
        }
    }
    if (dbg) for (int i = tid; i < Ne; i += M2_T) dbg[(size_t)Ne * HD + i] = x2[i];
    mbar_wait(bar, 0);
    M2_PHASE(2);

    // ---------------- D. pooling forward -------------------------------------------------------
    // B2[q] = [x2_gi, x2_gj, 1 - A, A] over the Ne-grid enumeration q; the L x L local grid selects
    // q = li (L-1) + lj - [lj > li]  (utils2.py:123-137, quirk Q3).  SP[li] = row sums, TP[lj] = column sums.
    const int nm1 = Ne - 1;
    const bool ident = Lb == Ne;
    if (ident) {
        float part = 0.f;
        for (int i = tid; i < Ne; i += M2_T) part += x2[i];
        const float X = mid2_block_sum(part, red);
        for (int i = warp; i < Ne; i += M2_NW) {            // row degrees
            const int c = __reduce_add_sync(0xffffffffu, lane < WPe ? __popc(ebits[i * WPe + lane]) : 0);
            if (lane == 0) SP[4 * i + 3] = (float)c;
        }
        for (int j = tid; j < Ne; j += M2_T) {              // column degrees
            const uint32_t* col = ebits + (j >> 5);
            const int sh = j & 31;
            int c = 0;
            for (int i = 0; i < Ne; ++i) c += (col[i * WPe] >> sh) & 1u;
            TP[4 * j + 3] = (float)c;
        }
        __syncthreads();
        for (int i = tid; i < Ne; i += M2_T) {
            const float xi = x2[i], fn = (float)nm1;
            SP[4 * i] = fn * xi; SP[4 * i + 1] = X - xi; SP[4 * i + 2] = fn - SP[4 * i + 3];
            TP[4 * i] = X - xi; TP[4 * i + 1] = fn * xi; TP[4 * i + 2] = fn - TP[4 * i + 3];
        }
    } else {
        const float inv = 1.f / (float)nm1;
        int nchunk = M2_T / Lb;
        nchunk = nchunk < 1 ? 1 : (nchunk > 4 ? 4 : nchunk);
        float* partR = scratch;                    // [nchunk][Lb][4]
        float* partC = scratch + 4 * 4 * Ne;
        for (int t = tid; t < nchunk * Lb; t += M2_T) {
            const int c = t / Lb, me = t - c * Lb;
            const int lo = (int)(((long long)c * Lb) / nchunk), hi = (int)(((long long)(c + 1) * Lb) / nchunk);
            float r0 = 0.f, r1 = 0.f, r3 = 0.f, q0 = 0.f, q1 = 0.f, q3 = 0.f;
            int cntr = 0, cntc = 0;
            for (int o = lo; o < hi; ++o) {
                if (o == me) continue;
                int gi, gj;
                unflat_pair(me * (Lb - 1) + o - (o > me), nm1, inv, gi, gj);          // row pass: li = me, lj = o
                r0 += x2[gi]; r1 += x2[gj]; r3 += (float)((ebits[gi * WPe + (gj >> 5)] >> (gj & 31)) & 1u); ++cntr;
                unflat_pair(o * (Lb - 1) + me - (me > o), nm1, inv, gi, gj);          // column pass: li = o, lj = me
                q0 += x2[gi]; q1 += x2[gj]; q3 += (float)((ebits[gi * WPe + (gj >> 5)] >> (gj & 31)) & 1u); ++cntc;
            }
            float* pr = partR + ((size_t)c * Lb + me) * 4;
            float* pc = partC + ((size_t)c * Lb + me) * 4;
            pr[0] = r0; pr[1] = r1; pr[2] = (float)cntr - r3; pr[3] = r3;
            pc[0] = q0; pc[1] = q1; pc[2] = (float)cntc - q3; pc[3] = q3;
        }
        __syncthreads();
        for (int idx = tid; idx < 4 * Lb; idx += M2_T) {
            float s = 0.f, t = 0.f;
            for (int c = 0; c < nchunk; ++c) { s += partR[(size_t)c * Lb * 4 + idx]; t += partC[(size_t)c * Lb * 4 + idx]; }
            SP[idx] = s; TP[idx] = t;
        }
    }
    __syncthreads();
    // segmented reduce by hunk id in ascending entity-line order
    for (int idx = tid; idx < Nc * 4; idx += M2_T) {
        const int c = idx >> 2, chn = idx & 3;
        float acc = 0.f;
        for (int i = 0; i < Lb; ++i)
            if (hm[i] == c) acc += SP[4 * i + chn] + TP[4 * i + chn];
        nb[idx] = acc;
        if (dbg) dbg[(size_t)Ne * 21 + idx] = acc;
    }
    __syncthreads();
    M2_PHASE(3);

    // ---------------- hunk-stage tables (union region) ---------------------------------------------
    const int T = Nc * HD;
    float* PH01 = uni;              // [Nc][KG][2][4]
    float* QH = uni + 2 * T;
    float* RS3 = uni + 3 * T;
    float* CS3 = uni + 4 * T;
    float* rr = uni + 5 * T;        // r, later GC
    float* cc = uni + 6 * T;
    float* PR01 = uni + 7 * T;      // [Nc][KG][2][4]; later dr (first T) and GR (second T)
    float* PC = uni + 9 * T;        // later dc
    float* RSm = uni + 10 * T;      // later RS3d
    float* CSm = uni + 11 * T;      // later CS3d
    float* dlt = uni + 12 * T;      // [Nc][CWT*32] dL/dlogit-difference per pair (training)
    constexpr int DW = CWT * 32;
    for (int idx = tid; idx < T; idx += M2_T) {
        const int c = idx / HD, k = idx - c * HD;
        float p = d1[k] + V1[8 * HD + k], q = 0.f;
#pragma unroll
        for (int chn = 0; chn < 4; ++chn) {
            p = fmaf(nb[4 * c + chn], V1[chn * HD + k], p);
            q = fmaf(nb[4 * c + chn], V1[(4 + chn) * HD + k], q);
        }
        const int pi = p01_idx(c, k);
        PH01[pi] = p; PH01[pi + 4] = p + Dh[k]; QH[idx] = q;
    }
    __syncthreads();

    // ---------------- E. hunk pair layer forward: row / column sums ---------------------------------
    {
        u64 Q[CWT][2], col[CWT][2];
#pragma unroll
        for (int sg = 0; sg < CWT; ++sg) {
            const int j = sg * 32 + lane;
            ulonglong2 q = make_ulonglong2(pk2(NEG_BIG, NEG_BIG), pk2(NEG_BIG, NEG_BIG));
            if (j < Nc) q = *reinterpret_cast<const ulonglong2*>(QH + j * HD + k0);
            Q[sg][0] = q.x; Q[sg][1] = q.y; col[sg][0] = 0ull; col[sg][1] = 0ull;
        }
        sweep2_fwd<CWT, false>(PH01, ybits, WPc, Nc, rg, M2_NRG, kg, Q, col, RS3, lane);
        combine_cols<CWT>(col, scratch, rg, M2_NRG, kg, lane);
        if (rg == 0) {
#pragma unroll
            for (int sg = 0; sg < CWT; ++sg) {
                const int j = sg * 32 + lane;
                if (j < Nc) *reinterpret_cast<ulonglong2*>(CS3 + j * HD + k0) = make_ulonglong2(col[sg][0], col[sg][1]);
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < T; idx += M2_T) {       // remove the diagonal pair (l = 0)
        const int c = idx / HD, k = idx - c * HD;
        const float d = fmaxf(PH01[p01_idx(c, k)] + QH[idx], 0.f);
        RS3[idx] -= d; CS3[idx] -= d;
        if (dbg) { dbg[(size_t)Ne * 21 + Nc * 4 + idx] = RS3[idx]; dbg[(size_t)Ne * 21 + Nc * 4 + T + idx] = CS3[idx]; }
    }
    __syncthreads();
    M2_PHASE(4);

    // ---------------- F. linear second layer on the sums + head tables (model_2.py:263-275, 311-315) --
    m2_mm20(rr, RS3, W2, b2, (float)(Nc - 1), Nc);
    m2_mm20(cc, CS3, W2, b2, (float)(Nc - 1), Nc);
    __syncthreads();
    for (int idx = tid; idx < T; idx += M2_T) {
        const int n = idx / HD, k = idx - n * HD;
        float p = g1b[k] + G1[k], q = 0.f;
#pragma unroll
        for (int m = 0; m < HD; ++m) {
            p = fmaf(rr[n * HD + m], G1[(2 + m) * HD + k], p);
            q = fmaf(cc[n * HD + m], G1[(2 + m) * HD + k], q);
        }
        const int pi = p01_idx(n, k);
        PR01[pi] = p; PR01[pi + 4] = p + Dg[k]; PC[idx] = q;
        if (dbg) { dbg[(size_t)Ne * 21 + Nc * 4 + 2 * T + idx] = p; dbg[(size_t)Ne * 21 + Nc * 4 + 3 * T + idx] = q; }
    }
    __syncthreads();
    M2_PHASE(5);

    // ---------------- G1. relation head: logits, softmax, CE (lanes = columns, all 20 channels) ---------
    float ce_acc = 0.f, d_acc = 0.f;
    {
        const size_t npair = (size_t)Nc * (Nc - 1);
        const float bd = gb2[1] - gb2[0], b20 = gb2[0];
        for (int cb = 0; cb < CWT; ++cb) {
            const int j = cb * 32 + lane;
            const bool ok = j < Nc;
            float Q[HD];
#pragma unroll
            for (int q4 = 0; q4 < 5; ++q4) {
                float4 v = make_float4(NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG);
                if (ok) v = *reinterpret_cast<const float4*>(PC + j * HD + 4 * q4);
                Q[4 * q4] = v.x; Q[4 * q4 + 1] = v.y; Q[4 * q4 + 2] = v.z; Q[4 * q4 + 3] = v.w;
            }
            for (int r = warp; r < Nc; r += M2_NW) {
                const bool valid = ok && j != r;
                const uint32_t bit = (ybits[r * WPc + cb] >> lane) & 1u;
                const bool lab = bit != 0u;
                const float* prow = PR01 + (size_t)r * PROW + bit * 4;
                float d = bd, l0 = b20;
#pragma unroll
                for (int q4 = 0; q4 < 5; ++q4) {
                    const float4 p = *reinterpret_cast<const float4*>(prow + q4 * 8);
                    const float h0 = fmaxf(p.x + Q[4 * q4], 0.f), h1 = fmaxf(p.y + Q[4 * q4 + 1], 0.f);
                    const float h2 = fmaxf(p.z + Q[4 * q4 + 2], 0.f), h3 = fmaxf(p.w + Q[4 * q4 + 3], 0.f);
                    d = fmaf(h0, gam[4 * q4], d); d = fmaf(h1, gam[4 * q4 + 1], d);
                    d = fmaf(h2, gam[4 * q4 + 2], d); d = fmaf(h3, gam[4 * q4 + 3], d);
                    if (LOGITS) {
                        l0 = fmaf(h0, G2[2 * (4 * q4)], l0); l0 = fmaf(h1, G2[2 * (4 * q4 + 1)], l0);
                        l0 = fmaf(h2, G2[2 * (4 * q4 + 2)], l0); l0 = fmaf(h3, G2[2 * (4 * q4 + 3)], l0);
                    }
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
                    const float dv = valid ? a.scale * (p1 - (lab ? 1.f : 0.f)) : 0.f;
                    d_acc += dv;
                    dlt[r * DW + j] = dv;
                }
            }
        }
    }
    {
        const float ce_tot = mid2_block_sum(ce_acc, red);
        if (tid == 0 && a.cep) a.cep[b] = ce_tot;
    }
    M2_PHASE(6);
    if (!TRAIN) return;

    // ---------------- G2. delta sums: RSm_i = sum_j m_ij dlt_ij, CSm_j, LSm (label-1 pairs) ----------------
    float* lsw = red + 64;                   // [M2_NRG][20]
    float* misc = red + 64 + M2_NW * HD;     // [0..19] LSm then LS4, [40] dsum
    {
        const float d_tot = mid2_block_sum(d_acc, red);      // leading __syncthreads orders the dlt stores
        if (tid == 0) misc[40] = d_tot;
        u64 Q[CWT][2], col[CWT][2], lsm[2] = {0ull, 0ull};
#pragma unroll
        for (int sg = 0; sg < CWT; ++sg) {
            const int j = sg * 32 + lane;
            ulonglong2 q = make_ulonglong2(pk2(NEG_BIG, NEG_BIG), pk2(NEG_BIG, NEG_BIG));
            if (j < Nc) q = *reinterpret_cast<const ulonglong2*>(PC + j * HD + k0);
            Q[sg][0] = q.x; Q[sg][1] = q.y; col[sg][0] = 0ull; col[sg][1] = 0ull;
        }
        const uint32_t lmask = 1u << lane;
        for (int r = rg; r < Nc; r += M2_NRG) {
            uint32_t w[CWT];
            load_words<CWT>(w, ybits + (size_t)r * WPc);
            const float* prow0 = PR01 + (size_t)r * PROW + kg * 8;
            const float* prow1 = prow0 + 4;
            const float* drow = dlt + (size_t)r * DW + lane;
            u64 rp0 = 0ull, rp1 = 0ull;
#pragma unroll
            for (int sg = 0; sg < CWT; ++sg) {
                const bool bit = (w[sg] & lmask) != 0u;
                const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(bit ? prow1 : prow0);
                const float dv = drow[sg * 32];
                const u64 d2 = pk2(dv, dv);
                const u64 v0 = gate2(add2(p.x, Q[sg][0]), d2), v1 = gate2(add2(p.y, Q[sg][1]), d2);
                col[sg][0] = add2(col[sg][0], v0); col[sg][1] = add2(col[sg][1], v1);
                rp0 = add2(rp0, v0); rp1 = add2(rp1, v1);
                const float lf = bit ? 1.f : 0.f;
                const u64 l2 = pk2(lf, lf);
                lsm[0] = fma2(l2, v0, lsm[0]); lsm[1] = fma2(l2, v1, lsm[1]);
            }
            const float tot = reduce4(rp0, rp1, lane);
            if ((lane & 7) == 0) RSm[r * HD + k0 + ch] = tot;
        }
        combine_cols<CWT>(col, scratch, rg, M2_NRG, kg, lane);
        if (rg == 0) {
#pragma unroll
            for (int sg = 0; sg < CWT; ++sg) {
                const int j = sg * 32 + lane;
                if (j < Nc) *reinterpret_cast<ulonglong2*>(CSm + j * HD + k0) = make_ulonglong2(col[sg][0], col[sg][1]);
            }
        }
        const float t = reduce4(lsm[0], lsm[1], lane);
        if ((lane & 7) == 0) lsw[rg * HD + k0 + ch] = t;
        __syncthreads();
        if (tid < HD) {
            float ls = 0.f;
            for (int w = 0; w < M2_NRG; ++w) ls += lsw[w * HD + tid];
            misc[tid] = ls;
        }
        __syncthreads();
    }
    M2_PHASE(7);

    // ---------------- H. head backward (node level) ------------------------------------------------------
    {
        // HS[k] = sum_pairs relu(pre)[k] * delta  via  relu(pre) = m * (PR0_i + l Dg + PC_j)
        if (tid < HD) {
            const int k = tid;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) {
                acc = fmaf(PR01[p01_idx(n, k)], RSm[n * HD + k], acc);
                acc = fmaf(PC[n * HD + k], CSm[n * HD + k], acc);
            }
            acc = fmaf(Dg[k], misc[k], acc);
            gp[po.scr_w2 + 2 * k + 1] = acc;
            gp[po.scr_w2 + 2 * k] = -acc;
        }
        __syncthreads();
        for (int idx = tid; idx < T; idx += M2_T) {       // RS4 = gam * RSm, CS4 = gam * CSm (in place)
            const int k = idx % HD;
            RSm[idx] *= gam[k]; CSm[idx] *= gam[k];
        }
        if (tid < HD) misc[tid] *= gam[tid];                     // LS4
        __syncthreads();
    }
    float* RS4 = RSm; float* CS4 = CSm;
    // scr_w1 rows 2.. : dG1e[m][k] = sum_n r[n][m] RS4[n][k] + c[n][m] CS4[n][k];  400 outputs
    for (int e = tid; e < 400 + HD; e += M2_T) {
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
    // dr = RS4 G1e^T -> PR01 buffer (first T) ; dc = CS4 G1e^T -> PC buffer
    float* dr = PR01; float* dc = PC;
    m2_mm20t(dr, RS4, G1 + 2 * HD, Nc);
    m2_mm20t(dc, CS4, G1 + 2 * HD, Nc);
    __syncthreads();
    // hnk_w2[q][m] = sum_n RS3[n][q] dr[n][m] + CS3[n][q] dc[n][m] ; hnk_b2[m] = (Nc-1) sum_n (dr+dc)[n][m]
    for (int e = tid; e < 400 + HD; e += M2_T) {
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
    m2_mm20t(GR, dr, W2, Nc);
    m2_mm20t(GC, dc, W2, Nc);
    float* RS3d = RSm; float* CS3d = CSm;
    __syncthreads();
    M2_PHASE(8);

    // ---------------- I. hunk pair layer backward sweep ------------------------------------------------------
    {
        u64 Q[CWT][2], GCr[CWT][2], col[CWT][2], ls3[2] = {0ull, 0ull};
#pragma unroll
        for (int sg = 0; sg < CWT; ++sg) {
            const int j = sg * 32 + lane;
            ulonglong2 q = make_ulonglong2(pk2(NEG_BIG, NEG_BIG), pk2(NEG_BIG, NEG_BIG)), g = make_ulonglong2(0ull, 0ull);
            if (j < Nc) {
                q = *reinterpret_cast<const ulonglong2*>(QH + j * HD + k0);
                g = *reinterpret_cast<const ulonglong2*>(GC + j * HD + k0);
            }
            Q[sg][0] = q.x; Q[sg][1] = q.y; GCr[sg][0] = g.x; GCr[sg][1] = g.y; col[sg][0] = 0ull; col[sg][1] = 0ull;
        }
        sweep2_bwd<CWT, false>(PH01, GR, ybits, WPc, Nc, rg, M2_NRG, kg, Q, GCr, col, ls3, RS3d, lane);
        combine_cols<CWT>(col, scratch, rg, M2_NRG, kg, lane);
        if (rg == 0) {
#pragma unroll
            for (int sg = 0; sg < CWT; ++sg) {
                const int j = sg * 32 + lane;
                if (j < Nc) *reinterpret_cast<ulonglong2*>(CS3d + j * HD + k0) = make_ulonglong2(col[sg][0], col[sg][1]);
            }
        }
        const float t = reduce4(ls3[0], ls3[1], lane);
        if ((lane & 7) == 0) lsw[rg * HD + k0 + ch] = t;
    }
    __syncthreads();
    for (int idx = tid; idx < T; idx += M2_T) {       // diagonal pair: l = 0
        const int c = idx / HD, k = idx - c * HD;
        const float d = (PH01[p01_idx(c, k)] + QH[idx]) > 0.f ? GR[idx] + GC[idx] : 0.f;
        RS3d[idx] -= d; CS3d[idx] -= d;
    }
    __syncthreads();
    M2_PHASE(9);

    // ---------------- J. hunk first-layer weights, d/dnb -----------------------------------------------------
    for (int e = tid; e < 8 * HD + HD; e += M2_T) {
        if (e < 8 * HD) {
            const int row = e / HD, k = e - row * HD, chn = row & 3;
            const float* src = row < 4 ? RS3d : CS3d;
            float acc = 0.f;
            for (int n = 0; n < Nc; ++n) acc = fmaf(nb[4 * n + chn], src[n * HD + k], acc);
            gp[po.hnk_w1 + e] = acc;
        } else {
            const int k = e - 8 * HD;
            float acc = 0.f, ls = 0.f;
            for (int n = 0; n < Nc; ++n) acc += RS3d[n * HD + k];
            for (int w = 0; w < M2_NRG; ++w) ls += lsw[w * HD + k];
            gp[po.hnk_b1 + k] = acc;
            gp[po.hnk_w1 + 9 * HD + k] = ls;
            gp[po.hnk_w1 + 8 * HD + k] = acc - ls;
        }
    }
    for (int idx = tid; idx < Nc * 4; idx += M2_T) {
        const int n = idx >> 2, chn = idx & 3;
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            v = fmaf(V1[chn * HD + k], RS3d[n * HD + k], v);
            v = fmaf(V1[(4 + chn) * HD + k], CS3d[n * HD + k], v);
        }
        dnb[idx] = v;
        if (dbg) dbg[(size_t)Ne * 21 + Nc * 4 + 4 * T + idx] = v;
    }
    __syncthreads();
    if (!a.ent) return;

    // ---------------- K. pooling backward ---------------------------------------------------------------------
    // dB2[q] = dnb[hunk(li)] + dnb[hunk(lj)] for q < L(L-1);  dx2[gi] += dB2[q][0], dx2[gj] += dB2[q][1]
    for (int idx = tid; idx < Ne * 4; idx += M2_T) {
        const int i = idx >> 2, chn = idx & 3;
        dl[idx] = (i < Lb && hm[i] >= 0) ? dnb[4 * hm[i] + chn] : 0.f;
    }
    __syncthreads();
    if (ident) {
        float p0 = 0.f, p1 = 0.f;
        for (int i = tid; i < Ne; i += M2_T) { p0 += dl[4 * i]; p1 += dl[4 * i + 1]; }
        const float D0 = mid2_block_sum(p0, red);
        const float D1 = mid2_block_sum(p1, red);
        const float fn = (float)nm1;
        for (int i = tid; i < Ne; i += M2_T) {
            const float a0 = dl[4 * i], a1 = dl[4 * i + 1];
            dx2[i] = (fmaf(fn, a0, D0 - a0)) + (fmaf(fn, a1, D1 - a1));
        }
    } else {
        const int lm1 = Lb - 1, qmax = Lb * lm1;
        const float invl = 1.f / (float)lm1;
        int nchunk = M2_T / Ne;
        nchunk = nchunk < 1 ? 1 : (nchunk > 4 ? 4 : nchunk);
        float* partR = scratch;                    // [nchunk][Ne]
        float* partC = scratch + 4 * Ne;
        for (int t = tid; t < nchunk * Ne; t += M2_T) {
            const int c = t / Ne, me = t - c * Ne;
            const int lo = (int)(((long long)c * Ne) / nchunk), hi = (int)(((long long)(c + 1) * Ne) / nchunk);
            float accr = 0.f, accc = 0.f;
            for (int o = lo; o < hi; ++o) {
                if (o == me) continue;
                int q = me * nm1 + o - (o > me);                 // row pass: gi = me, gj = o
                if (q < qmax) {
                    int li, lj;
                    unflat_pair(q, lm1, invl, li, lj);
                    accr += dl[4 * li] + dl[4 * lj];
                }
                q = o * nm1 + me - (me > o);                     // column pass: gi = o, gj = me
                if (q < qmax) {
                    int li, lj;
                    unflat_pair(q, lm1, invl, li, lj);
                    accc += dl[4 * li + 1] + dl[4 * lj + 1];
                }
            }
            partR[(size_t)c * Ne + me] = accr; partC[(size_t)c * Ne + me] = accc;
        }
        __syncthreads();
        for (int n = tid; n < Ne; n += M2_T) {
            float s = 0.f, t = 0.f;
            for (int c = 0; c < nchunk; ++c) { s += partR[(size_t)c * Ne + n]; t += partC[(size_t)c * Ne + n]; }
            dx2[n] = s + t;
        }
    }
    __syncthreads();
    if (dbg) for (int n = tid; n < Ne; n += M2_T) dbg[(size_t)Ne * 21 + Nc * 8 + 4 * T + n] = dx2[n];
    M2_PHASE(10);

    // ---------------- L. entity-state MLP backward (model_2.py:190-205) and W5/b5 (model_2.py:172-175) ------------
    {
        float* sS = uni; float* sE = sS + M2_CH * HD; float* sZp = sE + M2_CH * HD; float* sDz = sZp + M2_CH * HD;
        float* sDE = sDz + M2_CH * HD; float* sdu = sDE + M2_CH * HD;
        const float nb5 = 2.f * (float)(Ne - 1);
        constexpr int NREP = (881 + M2_T - 1) / M2_T;
        float accw[NREP];
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) accw[rep] = 0.f;
        for (int c0 = 0; c0 < Ne; c0 += M2_CH) {
            const int nn = min(M2_CH, Ne - c0);
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const size_t g = ((size_t)b * Ne + c0) * HD + idx;
                float v = a.RS1[g];
                for (int s = 0; s < nsl; ++s) v += a.CS1p[((size_t)b * a.SL + s) * Ne * HD + (size_t)c0 * HD + idx];
                sS[idx] = v;
            }
            __syncthreads();
            m2_mm20(sE, sS, W5, b5, nb5, nn);
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const int n = idx / HD, k = idx - n * HD;
                float acc = fmaf(xs[c0 + n], U1[k], c1[k]);
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(sE[n * HD + m], U1[(1 + m) * HD + k], acc);
                sZp[idx] = acc;
            }
            for (int n = tid; n < nn; n += M2_T) sdu[n] = x2[c0 + n] > 0.f ? dx2[c0 + n] : 0.f;
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const int n = idx / HD, k = idx - n * HD;
                sDz[idx] = sZp[idx] > 0.f ? sdu[n] * u2[k] : 0.f;
            }
            __syncthreads();
            for (int idx = tid; idx < nn * HD; idx += M2_T) {
                const int n = idx / HD, m = idx - n * HD;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < HD; ++k) acc = fmaf(U1[(1 + m) * HD + k], sDz[n * HD + k], acc);
                sDE[idx] = acc;
            }
            __syncthreads();
#pragma unroll
            for (int rep = 0; rep < NREP; ++rep) {
                const int e = tid + rep * M2_T;
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
                accw[rep] += acc;
            }
            for (int idx = tid; idx < nn * HD; idx += M2_T) {      // GE = dEbar W5^T
                const int n = idx / HD, q = idx - n * HD;
                float acc = 0.f;
#pragma unroll
                for (int m = 0; m < HD; ++m) acc = fmaf(W5[q * HD + m], sDE[n * HD + m], acc);
                a.GE[((size_t)b * Ne + c0) * HD + idx] = acc;
            }
            __syncthreads();
        }
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) {
            const int e = tid + rep * M2_T;
            const float v = accw[rep];
            if (e < 400) gp[po.nod_w1 + HD + e] = v;
            else if (e < 420) gp[po.nod_w1 + (e - 400)] = v;
            else if (e < 440) gp[po.nod_b1 + (e - 420)] = v;
            else if (e < 460) gp[po.nod_w2 + (e - 440)] = v;
            else if (e == 460) gp[po.nod_b2] = v;
            else if (e < 861) gp[po.ent_w5 + (e - 461)] = v;
            else if (e < 881) gp[po.ent_b5 + (e - 861)] = v;
        }
    }
    M2_PHASE(11);
}

}  // namespace hdgnn
