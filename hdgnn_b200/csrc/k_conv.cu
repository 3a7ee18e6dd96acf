// Legacy operator of the reference's model.py: symmetric adjacency normalisation
// fused with propagation, and the Chebyshev "map_conv" regulariser built on it.
//
//   normalize_adj   model.py:360-367   A_hat = D^-1/2 A^T D^-1/2,  D = rowsum(A) + 1e-3  (no self loop; the
//                                      transpose comes from tf.transpose(matmul(adj, D)) @ D)
//   chebyshev       model.py:335-391   T_0 = I, T_1 = (2 / 1.5)(I - A_hat) - I          (k = 2: the recurrence loop is empty)
//   map_conv        model.py:394-403   mean_b (x^T (softmax(theta)_0 T_0 + softmax(theta)_1 T_1) x)^2,  Ds = 1
//
// hdgnn_normalize_propagate is the general form  out = act(A_hat (H W) + b)  with switches for the north-star's
// A + I variant and for the un-transposed normalisation; default flags = the reference's formula.
//
// One CTA per commit.  The whole per-commit adjacency tile (N x pitch bytes) is staged in shared memory by ONE
// 1-D TMA bulk copy; the degree scan runs while the tile is resident and packs the rows into bitmaps, the
// bitmap is transposed with ballots when the reference's A^T form is asked for, and the propagation walks the
// set bits (lanes = output channels).  HBM traffic = the adjacency bytes once + H in + out.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/hdgnn.h"
#include "common.cuh"

namespace hdgnn {

constexpr int CONV_T = 256;
constexpr int CONV_MAXD = 32;       // d_in, d_out <= 32

struct ConvArgs {
    int B, N, pitch, d_in, d_out, flags;
    float eps, lam_max;
    const uint8_t* adj; const float* H; const float* W; const float* bias;
    float* out; float* dinv_out;
    const float* x; const float* theta; float* per_commit;       // map_conv
};

__host__ __device__ inline int conv_words(int N) { return (N + 31) / 32; }
__host__ __device__ inline size_t conv_smem_bytes(int N, int pitch, int d_out) {
    const int WP = conv_words(N), NP = WP * 32;
    size_t off = 16 + (size_t)round_up(N * pitch, 128);     // mbarrier, byte tile
    off += (size_t)2 * NP * WP * 4;                          // row bitmap, column bitmap (padded to 32 rows)
    off += (size_t)N * 4;                                    // dinv
    off += (size_t)N * (d_out > 0 ? d_out : 1) * 4;          // Hs (or x scaled)
    off += (size_t)(CONV_MAXD * CONV_MAXD + CONV_MAXD + 64) * 4;
    return off + 16;
}

__device__ __forceinline__ uint32_t nz4(uint32_t v) {
    const uint32_t t = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
    return ((t >> 7) | (t >> 14) | (t >> 21) | (t >> 28)) & 0xfu;
}

// Stages the tile, builds rbits[j][w] (bit i of word i>>5 = A[j][i] != 0, diagonal cleared), the degrees and
// dinv[j] = (deg_j + [self loop] + eps)^-1/2, and (if `transpose`) cbits[i][w] = rbits^T.  Returns pointers.
struct ConvSmem { uint8_t* tile; uint32_t* rbits; uint32_t* cbits; float* dinv; float* Hs; float* Ws; float* red; uint64_t* bar; };

__device__ __forceinline__ ConvSmem conv_carve(unsigned char* smem, int N, int pitch, int d_out) {
    const int WP = conv_words(N), NP = WP * 32;
    ConvSmem s;
    s.bar = reinterpret_cast<uint64_t*>(smem);
    s.tile = smem + 16;
    s.rbits = reinterpret_cast<uint32_t*>(smem + 16 + round_up(N * pitch, 128));
    s.cbits = s.rbits + (size_t)NP * WP;
    s.dinv = reinterpret_cast<float*>(s.cbits + (size_t)NP * WP);
    s.Hs = s.dinv + N;
    s.Ws = s.Hs + (size_t)N * (d_out > 0 ? d_out : 1);
    s.red = s.Ws + CONV_MAXD * CONV_MAXD + CONV_MAXD;
    return s;
}

__device__ __forceinline__ void conv_degree_scan(const ConvSmem& s, int N, int pitch, bool self_loop, float eps, bool transpose) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, WP = conv_words(N), NP = WP * 32;
    // rows: 32 lanes x 16 bytes cover 512 columns; one row per warp iteration
    for (int j = warp; j < NP; j += CONV_T / 32) {
        uint32_t m = 0u;
        if (j < N && lane * 16 < pitch) {
            const uint4 v = *reinterpret_cast<const uint4*>(s.tile + (size_t)j * pitch + lane * 16);
            m = nz4(v.x) | (nz4(v.y) << 4) | (nz4(v.z) << 8) | (nz4(v.w) << 12);
        }
        const uint32_t hi = __shfl_down_sync(0xffffffffu, m, 1);
        uint32_t w = 0u;
        const int sg = lane >> 1;
        if ((lane & 1) == 0 && sg < WP) {
            w = m | (hi << 16);
            const int c0 = sg * 32;
            if (c0 + 32 > N) w &= (c0 >= N) ? 0u : (0xffffffffu >> (c0 + 32 - N));
            if ((j >> 5) == sg) w &= ~(1u << (j & 31));
            s.rbits[(size_t)j * WP + sg] = w;
        }
        const int deg = __reduce_add_sync(0xffffffffu, __popc(w));
        if (lane == 0 && j < N) s.dinv[j] = 1.f / sqrtf((float)deg + (self_loop ? 1.f : 0.f) + eps);
    }
    __syncthreads();
    if (transpose) {
        // 32 x 32 bit-block transposes: lane l holds row (jw*32 + l)'s word iw; ballot over bit c gives column c
        for (int blk = warp; blk < WP * WP; blk += CONV_T / 32) {
            const int jw = blk / WP, iw = blk - jw * WP;
            const uint32_t word = s.rbits[(size_t)(jw * 32 + lane) * WP + iw];
            uint32_t mine = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const uint32_t col = __ballot_sync(0xffffffffu, (word >> c) & 1u);
                if (lane == c) mine = col;
            }
            s.cbits[(size_t)(iw * 32 + lane) * WP + jw] = mine;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CONV_T) propagate_kernel(const ConvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, pitch = a.pitch, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d_in = a.d_in, d_out = a.d_out, WP = conv_words(N);
    const ConvSmem s = conv_carve(smem, N, pitch, d_out);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(s.bar, (uint32_t)N * pitch);
        bulk_g2s(s.tile, a.adj + (size_t)b * N * pitch, (uint32_t)N * pitch, s.bar);
    }
    // H W while the tile is in flight
    if (a.W) for (int i = tid; i < d_in * d_out; i += CONV_T) s.Ws[i] = a.W[i];
    __syncthreads();
    const float* Hb = a.H + (size_t)b * N * d_in;
    for (int idx = tid; idx < N * d_out; idx += CONV_T) {
        const int j = idx / d_out, c = idx - j * d_out;
        float acc = 0.f;
        if (a.W) { for (int k = 0; k < d_in; ++k) acc = fmaf(Hb[(size_t)j * d_in + k], s.Ws[k * d_out + c], acc); }
        else acc = Hb[(size_t)j * d_in + c];
        s.Hs[idx] = acc;
    }
    mbar_wait(s.bar, 0);
    __syncthreads();
    const bool self_loop = a.flags & HDGNN_P_SELF_LOOP, transpose = !(a.flags & HDGNN_P_NO_TRANSPOSE);
    conv_degree_scan(s, N, pitch, self_loop, a.eps, transpose);
    for (int idx = tid; idx < N * d_out; idx += CONV_T) s.Hs[idx] *= s.dinv[idx / d_out];
    if (a.dinv_out) for (int j = tid; j < N; j += CONV_T) a.dinv_out[(size_t)b * N + j] = s.dinv[j];
    __syncthreads();
    // out_i = act(dinv_i * sum_{j in nbr(i)} Hs_j + b): lanes = channels, one output row per warp iteration
    const uint32_t* nb = transpose ? s.cbits : s.rbits;
    for (int i = warp; i < N; i += CONV_T / 32) {
        float acc = (self_loop && lane < d_out) ? s.Hs[(size_t)i * d_out + lane] : 0.f;
        for (int w = 0; w < WP; ++w) {
            uint32_t bits = nb[(size_t)i * WP + w];
            while (bits) {
                const int j = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                if (lane < d_out) acc += s.Hs[(size_t)j * d_out + lane];
            }
        }
        if (lane < d_out) {
            float v = fmaf(s.dinv[i], acc, a.bias ? a.bias[lane] : 0.f);
            if (a.flags & HDGNN_P_RELU) v = fmaxf(v, 0.f);
            a.out[((size_t)b * N + i) * d_out + lane] = v;
        }
    }
}

// ---- tcgen05 variant of the propagation ---------------------------------------------------------------------
// The per-commit product  (A or A^T, N x N, entries 0/1)  x  (D^-1/2 H W, N x d_out)  is a dense contraction:
// the adjacency is exact in bf16 and the scaled features are split into three bf16 terms (hi + mid + lo carries
// 24 mantissa bits), so one tcgen05.mma chain with fp32 accumulation in TMEM reproduces the fp32 result whatever
// the edge density.  Operands are written to shared memory by the CTA itself in the canonical no-swizzle K-major
// layout ([K/8][rows][8] bf16: 8 x 16-byte core matrices, LBO = rows * 16 B, SBO = 128 B); one thread issues
// the MMAs (M = 128 per tile, N = 4 d_out rounded to 16, K = 16 per instruction), tcgen05.commit signals an
// mbarrier and the eight warps read their 32 TMEM lanes back with tcgen05.ld for the epilogue.
constexpr int TC_T = 512;

__host__ __device__ inline int tc_np(int d_out) { return round_up(4 * d_out, 16); }
__host__ __device__ inline int tc_mp(int N) { return N <= 128 ? 128 : 256; }
__host__ __device__ inline size_t tc_smem_bytes(int N, int pitch, int d_in, int d_out) {
    const int WP = conv_words(N), NP32 = WP * 32, KP = round_up(N, 16);
    const size_t tile = (size_t)round_up(N * pitch, 128), sa = (size_t)tc_mp(N) * KP * 2;
    size_t off = 32 + (tile > sa ? tile : sa);                 // mbarrier + tmem slot | byte tile, later the A operand
    off += (size_t)tc_np(d_out) * KP * 2;                      // B operand
    off += (size_t)NP32 * WP * 4;                              // row bitmap
    off += (size_t)round_up(N, 4) * 4 + (size_t)round_up(N * d_out, 4) * 4;     // dinv, Hs
    off += (size_t)(CONV_MAXD * CONV_MAXD + CONV_MAXD) * 4;    // W
    off += (size_t)round_up(N * d_in, 4) * 4;                  // H tile (TMA)
    return off + 128;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // cute/arch/mma_sm100_desc.hpp UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
// UMMA::InstrDescriptor: D = F32 [4,6), A = B = BF16 [7,10) [10,13), K-major, N >> 3 [17,23), M >> 4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(TC_T, 1) propagate_tc_kernel(const ConvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, pitch = a.pitch, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d_in = a.d_in, d_out = a.d_out, WP = conv_words(N), NP32 = WP * 32, KP = round_up(N, 16);
    const int MP = tc_mp(N), NPc = tc_np(d_out), MT = MP / 128;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    uint8_t* tile = smem + 32;
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem + 32);                    // aliases the byte tile
    const size_t tile_b = (size_t)round_up(N * pitch, 128), sa_b = (size_t)MP * KP * 2;
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + 32 + (tile_b > sa_b ? tile_b : sa_b));
    uint32_t* rbits = reinterpret_cast<uint32_t*>(sB + (size_t)NPc * KP);
    float* dinv = reinterpret_cast<float*>(rbits + (size_t)NP32 * WP);
    float* Hs = dinv + round_up(N, 4);
    float* Ws = Hs + (size_t)round_up(N * d_out, 4);
    float* Hin = Ws + CONV_MAXD * CONV_MAXD + CONV_MAXD;        // 16-byte aligned: every block above is a multiple of 4 floats
    int tcols = 32;
    while (tcols < MT * NPc) tcols <<= 1;
    if (tid == 0) {
        mbar_init(bar_tma, 1); mbar_init(bar_mma, 1);
        fence_mbar_init();
        const uint32_t hbytes = (uint32_t)N * d_in * 4;
        const bool htma = (hbytes & 15u) == 0 && ((((size_t)b * N * d_in * 4) & 15) == 0) && (((uintptr_t)a.H & 15) == 0);
        mbar_arrive_expect_tx(bar_tma, (uint32_t)N * pitch + (htma ? hbytes : 0u));
        bulk_g2s(tile, a.adj + (size_t)b * N * pitch, (uint32_t)N * pitch, bar_tma);
        if (htma) bulk_g2s(Hin, a.H + (size_t)b * N * d_in, hbytes, bar_tma);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.W) for (int i = tid; i < d_in * d_out; i += TC_T) Ws[i] = a.W[i];
    {
        const uint32_t hbytes = (uint32_t)N * d_in * 4;
        const bool htma = (hbytes & 15u) == 0 && ((((size_t)b * N * d_in * 4) & 15) == 0) && (((uintptr_t)a.H & 15) == 0);
        if (!htma) {            // odd shapes: plain coalesced copy of the feature tile
            const float* Hb = a.H + (size_t)b * N * d_in;
            for (int i = tid; i < N * d_in; i += TC_T) Hin[i] = Hb[i];
        }
    }
    mbar_wait(bar_tma, 0);
    __syncthreads();
    for (int idx = tid; idx < N * d_out; idx += TC_T) {         // H W from shared memory
        const int j = idx / d_out, c = idx - j * d_out;
        float acc = 0.f;
        if (a.W) { for (int k = 0; k < d_in; ++k) acc = fmaf(Hin[j * d_in + k], Ws[k * d_out + c], acc); }
        else acc = Hin[j * d_in + c];
        Hs[idx] = acc;
    }
    const bool self_loop = a.flags & HDGNN_P_SELF_LOOP, transpose = !(a.flags & HDGNN_P_NO_TRANSPOSE);
    // degree scan + row bitmap while the byte tile is resident (same as conv_degree_scan, no transpose needed)
    for (int j = warp; j < NP32; j += TC_T / 32) {
        uint32_t m = 0u;
        if (j < N && lane * 16 < pitch) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)j * pitch + lane * 16);
            m = nz4(v.x) | (nz4(v.y) << 4) | (nz4(v.z) << 8) | (nz4(v.w) << 12);
        }
        const uint32_t hi = __shfl_down_sync(0xffffffffu, m, 1);
        uint32_t w = 0u;
        const int sg = lane >> 1;
        if ((lane & 1) == 0 && sg < WP) {
            w = m | (hi << 16);
            const int c0 = sg * 32;
            if (c0 + 32 > N) w &= (c0 >= N) ? 0u : (0xffffffffu >> (c0 + 32 - N));
            if ((j >> 5) == sg) w &= ~(1u << (j & 31));
            rbits[(size_t)j * WP + sg] = w;
        }
        const int deg = __reduce_add_sync(0xffffffffu, __popc(w));
        if (lane == 0 && j < N) dinv[j] = 1.f / sqrtf((float)deg + (self_loop ? 1.f : 0.f) + a.eps);
    }
    __syncthreads();                    // the byte tile is dead from here on: sA overwrites it
    if (a.dinv_out) for (int j = tid; j < N; j += TC_T) a.dinv_out[(size_t)b * N + j] = dinv[j];
    // A operand: A_mma[m = i][k = j] = adj[j][i] (reference form) or adj[i][j]; + I with the self loop; zero padding
    const uint32_t one = 0x3F80u;       // bf16 1.0
    for (int t = tid; t < MP * (KP >> 3); t += TC_T) {
        const int kk = t / MP, m = t - kk * MP;
        uint32_t bits8 = 0u;
        if (m < N) {
            if (transpose) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = kk * 8 + u;
                    bits8 |= ((rbits[(size_t)j * WP + (m >> 5)] >> (m & 31)) & 1u) << u;     // rows j >= N are zero (padded bitmap)
                }
            } else {
                bits8 = (rbits[(size_t)m * WP + (kk >> 2)] >> ((kk & 3) * 8)) & 0xffu;
            }
            if (self_loop && (m >> 3) == kk) bits8 |= 1u << (m & 7);
        }
        uint4 v;
        v.x = ((bits8 & 1u) ? one : 0u) | ((bits8 & 2u) ? one << 16 : 0u);
        v.y = ((bits8 & 4u) ? one : 0u) | ((bits8 & 8u) ? one << 16 : 0u);
        v.z = ((bits8 & 16u) ? one : 0u) | ((bits8 & 32u) ? one << 16 : 0u);
        v.w = ((bits8 & 64u) ? one : 0u) | ((bits8 & 128u) ? one << 16 : 0u);
        *reinterpret_cast<uint4*>(sA + ((size_t)kk * MP + m) * 8) = v;
    }
    // B operand: B_mma[n = 4 c + s][k = j] = s-th bf16 term of dinv_j * Hs[j][c] (s = 0, 1, 2), zero elsewhere
    for (int t = tid; t < (NPc >> 2) * (KP >> 3); t += TC_T) {
        const int kk = t / (NPc >> 2), c = t - kk * (NPc >> 2);
        __nv_bfloat16 o0[8], o1[8], o2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = kk * 8 + u;
            const float v = (c < d_out && j < N) ? Hs[(size_t)j * d_out + c] * dinv[j] : 0.f;
            o0[u] = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(o0[u]);
            o1[u] = __float2bfloat16_rn(r1);
            o2[u] = __float2bfloat16_rn(r1 - __bfloat162float(o1[u]));
        }
        uint4* dst = reinterpret_cast<uint4*>(sB + ((size_t)kk * NPc + 4 * c) * 8);
        dst[0] = *reinterpret_cast<const uint4*>(o0); dst[1] = *reinterpret_cast<const uint4*>(o1);
        dst[2] = *reinterpret_cast<const uint4*>(o2); dst[3] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();                // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, NPc);
        for (int mt = 0; mt < MT; ++mt) {
            for (int ks = 0; ks < (KP >> 4); ++ks) {
                const uint64_t da = umma_desc(smem_u32(sA) + (uint32_t)(ks * 2 * MP + mt * 128) * 16, (uint32_t)MP * 16, 128);
                const uint64_t db = umma_desc(smem_u32(sB) + (uint32_t)(ks * 2 * NPc) * 16, (uint32_t)NPc * 16, 128);
                const uint32_t accum = ks > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)(mt * NPc)), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
    }
    mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 of M tile w / 4
    if (warp / 4 < MT) {
        const int mt = warp >> 2, i = mt * 128 + (warp & 3) * 32 + lane;
        const float di = i < N ? dinv[i] : 0.f;
        for (int c0 = 0; c0 < d_out; c0 += 4) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * NPc + 4 * c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (i < N) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = c0 + q;
                    if (c < d_out) {
                        const float acc = (__uint_as_float(r[4 * q]) + __uint_as_float(r[4 * q + 1])) + __uint_as_float(r[4 * q + 2]);
                        float v = fmaf(di, acc, a.bias ? a.bias[c] : 0.f);
                        if (a.flags & HDGNN_P_RELU) v = fmaxf(v, 0.f);
                        a.out[((size_t)b * N + i) * d_out + c] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tcols) : "memory");
}

// per commit: q_b = x^T (t0 x + t1 ((2/lam)(x - A_hat x) - x)),  t = softmax(theta)
__global__ void __launch_bounds__(CONV_T) map_conv_kernel(const ConvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, pitch = a.pitch, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WP = conv_words(N);
    const ConvSmem s = conv_carve(smem, N, pitch, 1);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(s.bar, (uint32_t)N * pitch);
        bulk_g2s(s.tile, a.adj + (size_t)b * N * pitch, (uint32_t)N * pitch, s.bar);
    }
    __syncthreads();
    mbar_wait(s.bar, 0);
    const bool self_loop = a.flags & HDGNN_P_SELF_LOOP, transpose = !(a.flags & HDGNN_P_NO_TRANSPOSE);
    conv_degree_scan(s, N, pitch, self_loop, a.eps, false);
    const float* xb = a.x + (size_t)b * N;
    for (int j = tid; j < N; j += CONV_T) s.Hs[j] = xb[j] * s.dinv[j];
    __syncthreads();
    // (A_hat x)_i = dinv_i sum_j A[j][i] xs_j (reference form) or dinv_i sum_j A[i][j] xs_j
    float part = 0.f;
    const float th0 = a.theta[0], th1 = a.theta[1], mx = fmaxf(th0, th1);
    const float e0 = expf(th0 - mx), e1 = expf(th1 - mx), t0 = e0 / (e0 + e1), t1 = e1 / (e0 + e1);
    if (transpose) {
        // thread per column i: walk the rows, one broadcast word per row
        for (int i = tid; i < N; i += CONV_T) {
            const int w = i >> 5, sh = i & 31;
            float acc = self_loop ? s.Hs[i] : 0.f;
            for (int j = 0; j < N; ++j) acc += ((s.rbits[(size_t)j * WP + w] >> sh) & 1u) ? s.Hs[j] : 0.f;
            const float xi = xb[i], ax = s.dinv[i] * acc;
            part += xi * fmaf(t1, (2.f / a.lam_max) * (xi - ax) - xi, t0 * xi);
        }
    } else {
        for (int i = tid; i < N; i += CONV_T) {
            float acc = self_loop ? s.Hs[i] : 0.f;
            for (int w = 0; w < WP; ++w) {
                uint32_t bits = s.rbits[(size_t)i * WP + w];
                while (bits) { acc += s.Hs[w * 32 + __ffs(bits) - 1]; bits &= bits - 1; }
            }
            const float xi = xb[i], ax = s.dinv[i] * acc;
            part += xi * fmaf(t1, (2.f / a.lam_max) * (xi - ax) - xi, t0 * xi);
        }
    }
    const float q = block_sum(part, s.red);
    if (tid == 0) a.per_commit[b] = q * q;
    (void)lane; (void)warp;
}

__global__ void mean_kernel(const float* v, int n, float* out) {
    __shared__ float scratch[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
    const float t = block_sum(acc, scratch);
    if (threadIdx.x == 0) *out = t / (float)n;
}

}  // namespace hdgnn

using namespace hdgnn;

static int conv_check(int B, int N, int pitch, const void* adj) {
    if (B < 1 || N < 2 || N > HDGNN_MAX_N || !adj) return HDGNN_E_INVALID;
    if (pitch < N || (pitch & 15) || ((uintptr_t)adj & 15)) return HDGNN_E_INVALID;
    return HDGNN_OK;
}

extern "C" int hdgnn_normalize_propagate(int B, int N, const uint8_t* adj, int pitch, const float* H, int d_in, const float* W,
                                         const float* bias, int d_out, float eps, int flags, float* out, float* dinv_out,
                                         void* stream) {
    int rc = conv_check(B, N, pitch, adj);
    if (rc) return rc;
    if (!H || !out || d_in < 1 || d_in > CONV_MAXD || d_out < 1 || d_out > CONV_MAXD) return HDGNN_E_INVALID;
    if (!W && d_in != d_out) return HDGNN_E_INVALID;
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return HDGNN_E_CUDA;
    ConvArgs a{};
    a.B = B; a.N = N; a.pitch = pitch; a.d_in = d_in; a.d_out = d_out; a.flags = flags; a.eps = eps;
    a.adj = adj; a.H = H; a.W = W; a.bias = bias; a.out = out; a.dinv_out = dinv_out;
    // tensor-core path (tcgen05): N <= 256 and the operands fit one SM; else the set-bit walk on the CUDA cores
    const size_t smem_tc = tc_smem_bytes(N, pitch, d_in, d_out);
    if (!(flags & HDGNN_P_NO_TENSOR) && N <= 256 && smem_tc <= (size_t)optin) {
        if (cudaFuncSetAttribute(propagate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) return HDGNN_E_CUDA;
        propagate_tc_kernel<<<B, TC_T, smem_tc, (cudaStream_t)stream>>>(a);
        return cudaGetLastError() == cudaSuccess ? HDGNN_OK : HDGNN_E_CUDA;
    }
    const size_t smem = conv_smem_bytes(N, pitch, d_out);
    if (smem > (size_t)optin) return HDGNN_E_UNSUPPORTED;       // the per-commit tile must fit one SM
    if (cudaFuncSetAttribute(propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) return HDGNN_E_CUDA;
    propagate_kernel<<<B, CONV_T, smem, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? HDGNN_OK : HDGNN_E_CUDA;
}

extern "C" int hdgnn_map_conv(int B, int N, const uint8_t* adj, int pitch, const float* x, const float* theta, float lam_max,
                              float eps, int flags, float* per_commit, float* loss, void* stream) {
    int rc = conv_check(B, N, pitch, adj);
    if (rc) return rc;
    if (!x || !theta || !per_commit || lam_max <= 0.f) return HDGNN_E_INVALID;
    const size_t smem = conv_smem_bytes(N, pitch, 1);
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return HDGNN_E_CUDA;
    if (smem > (size_t)optin) return HDGNN_E_UNSUPPORTED;
    if (cudaFuncSetAttribute(map_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) return HDGNN_E_CUDA;
    ConvArgs a{};
    a.B = B; a.N = N; a.pitch = pitch; a.flags = flags; a.eps = eps; a.lam_max = lam_max;
    a.adj = adj; a.x = x; a.theta = theta; a.per_commit = per_commit;
    map_conv_kernel<<<B, CONV_T, smem, (cudaStream_t)stream>>>(a);
    if (loss) mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(per_commit, B, loss);
    return cudaGetLastError() == cudaSuccess ? HDGNN_OK : HDGNN_E_CUDA;
}
