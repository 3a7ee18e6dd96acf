// k_io: the two data formats either side of the hot path, both HBM-bound byte/integer work.
//
//  * compact_from_raw_kernel  -- utils2.py:29-47 and the int() indexing of utils2.py:82,105 on the device: the arrays as
//    stored in CAdjs_{step}.npy / CHunkAdjs_{step}.npy (float64 or float32, diagonal = node attribute) become the u8
//    label grid + f32 diagonal the hot path reads.  8 (or 4) bytes in, 1 byte out per element.
//  * eval_counts_kernel       -- EvaluationFuncs.py:27-37 (top_ACC), :92-117 (prec / recall / f1, with and without the
//    ceil-on-channel-0 quirk) and :119-153 (AUC) as integer counters per commit, read straight from the probs the
//    relation head wrote; the (N,2,Ncr) device->host copy and the Python double loops disappear.
//
// Everything here is integer counting => bit-exact and order independent (integer atomics).
#include "../../include/hdgnn.h"
#include "common.cuh"

namespace hdgnn {

// ---------------------------------------------------------------------------------------------------------------
// raw (N,n,n) -> grid u8 (N,n,pitch), diag f32 (N,n).  One warp per ROWS_PER_WARP consecutive rows; lanes own the columns
// lane + 32 k, so every load instruction of a warp covers 256 (f64) contiguous bytes.  All loads of a row are in
// flight before the first convert.
// int(v) of utils2.py:82,105 truncates toward zero and indexes a size-2 axis: {-2,-1,0,1} are legal (python negative
// indices), anything else raises IndexError there -> *err = 1 here.  NaN fails the range test as well.
template <typename T, int KMAX>
__global__ void __launch_bounds__(256) compact_from_raw_kernel(const T* __restrict__ raw, int rows, int n, uint8_t* __restrict__ grid,
                                                               int pitch, float* __restrict__ diag, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    int bad = 0;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        const int i = row % n;                              // row index inside its commit
        const T* src = raw + (size_t)row * n;
        T v[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int j = lane + 32 * k;
            v[k] = j < n ? __ldg(src + j) : T(0);
        }
        uint8_t* dst = grid + (size_t)row * pitch;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int j = lane + 32 * k;
            if (j >= pitch) continue;
            uint8_t o = 0;
            if (j < n) {
                if (j == i) {
                    if (diag) diag[row] = (float)v[k];      // utils2.py:31-36: x_i = A_ii (float64 -> float32, round to nearest)
                } else {
                    const T t = v[k];
                    if (!(t > T(-3) && t < T(2))) bad = 1;
                    o = (uint8_t)(((int)t) & 1);            // -1 -> channel 1, -2 -> channel 0 (python indexing from the end)
                }
            }
            dst[j] = o;
        }
    }
    if (bad && err) atomicOr(err, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// counts (B,8) int64: [0] arg-max hits, [1..3] tp fp fn of the reference's quirk form (y_true = ceil(label ch 0) = 1 - Y,
// y_pred = ceil(prob ch 0) = [prob0 > 0]), [4..6] tp fp fn of the conventional form (y_true = Y, y_pred = [p1 > p0]),
// [7] number of related pairs.  auc (B,2) int64: Mann-Whitney numerators 2 #{(pos,neg): s_neg < s_pos} + #{s_neg == s_pos}
// with [0] the reference's scoring (EvaluationFuncs.py:128-143: a pair is scored with the probability of the channel
// its label does NOT have) and [1] the conventional one (score = p1).  AUC = auc / (2 npos nneg).
constexpr int EV_T = 256;
constexpr int EV_PER = 8;                   // pairs per thread
constexpr int EV_CHUNK = EV_T * EV_PER;     // pairs per CTA
constexpr int EV_TILE = 2048;               // positives staged per round

__global__ void __launch_bounds__(EV_T) eval_counts_kernel(int Nc, const float* __restrict__ probs, const uint8_t* __restrict__ Y,
                                                           int y_pitch, unsigned long long* __restrict__ counts,
                                                           unsigned long long* __restrict__ auc, int auc_first) {
    __shared__ float posq[EV_TILE], posc[EV_TILE];
    __shared__ int npos_s;
    __shared__ unsigned long long red[8];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int nm1 = Nc - 1, Ncr = Nc * nm1;
    const float inv = 1.f / (float)nm1;
    const float* p0 = probs + (size_t)b * 2 * Ncr;
    const float* p1 = p0 + Ncr;
    const uint8_t* Yb = Y + (size_t)b * Nc * y_pitch;
    if (tid < 8) red[tid] = 0ull;
    __syncthreads();

    float sneg[EV_PER];                      // this thread's pairs as negatives (score p1 in both forms); +inf = not a negative
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < EV_PER; ++k) {
        const int p = blockIdx.x * EV_CHUNK + k * EV_T + tid;
        sneg[k] = __int_as_float(0x7f800000);
        if (p < Ncr) {
            int s, t;
            unflat_pair(p, nm1, inv, s, t);
            const int y = Yb[(size_t)s * y_pitch + t] != 0;
            const float a = p0[p], d = p1[p];
            const int am = d > a;                             // np.argmax: ties -> channel 0
            c[0] += am == y;
            const int qt = !y, qp = a > 0.f;                  // quirk form scores channel 0
            c[1] += qt & qp; c[2] += (!qt) & qp; c[3] += qt & (!qp);
            c[4] += y & am; c[5] += (!y) & am; c[6] += y & (!am);
            c[7] += y;
            if (!y) sneg[k] = d;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int w = __reduce_add_sync(0xffffffffu, c[q]);
        if (lane == 0 && w) atomicAdd(&red[q], (unsigned long long)w);
    }
    __syncthreads();
    if (tid < 8 && red[tid]) atomicAdd(&counts[(size_t)b * 8 + tid], red[tid]);
    if (!auc || b < auc_first) return;

    // every CTA of the commit walks ALL pairs of the commit for the positives, EV_TILE candidates at a time
    unsigned long long nq = 0ull, ncv = 0ull;
    for (int base = 0; base < Ncr; base += EV_TILE) {
        __syncthreads();
        if (tid == 0) npos_s = 0;
        __syncthreads();
        for (int p = base + tid; p < min(base + EV_TILE, Ncr); p += EV_T) {
            int s, t;
            unflat_pair(p, nm1, inv, s, t);
            if (Yb[(size_t)s * y_pitch + t] != 0) {
                const int slot = atomicAdd(&npos_s, 1);       // order is irrelevant: the counters below are integers
                posq[slot] = p0[p]; posc[slot] = p1[p];
            }
        }
        __syncthreads();
        const int np = npos_s;
        unsigned int lq = 0u, lc = 0u;                        // <= 2 * EV_TILE * EV_PER per round
        for (int i = 0; i < np; ++i) {
            const float sq = posq[i], sc = posc[i];
#pragma unroll
            for (int k = 0; k < EV_PER; ++k) {
                lq += (sneg[k] < sq ? 2u : 0u) + (sneg[k] == sq ? 1u : 0u);
                lc += (sneg[k] < sc ? 2u : 0u) + (sneg[k] == sc ? 1u : 0u);
            }
        }
        nq += lq; ncv += lc;
    }
    nq = __reduce_add_sync(0xffffffffu, (unsigned int)(nq & 0xffffffffu)) + ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)(nq >> 32)) << 32);
    ncv = __reduce_add_sync(0xffffffffu, (unsigned int)(ncv & 0xffffffffu)) + ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)(ncv >> 32)) << 32);
    if (lane == 0) {
        if (nq) atomicAdd(&auc[(size_t)b * 2], nq);
        if (ncv) atomicAdd(&auc[(size_t)b * 2 + 1], ncv);
    }
}

}  // namespace hdgnn

using namespace hdgnn;

extern "C" int hdgnn_compact_from_raw(int N, int n, const void* raw, int raw_is_f64, uint8_t* grid, int pitch, float* diag,
                                      int32_t* err, void* stream) {
    if (N < 1 || n < 2 || n > HDGNN_MAX_N || !raw || !grid) return HDGNN_E_INVALID;
    if (pitch < n || (pitch & 15) || pitch > 512 + 16) return HDGNN_E_INVALID;
    const long long rows = (long long)N * n;
    if (rows > 0x7fffffffLL) return HDGNN_E_INVALID;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return HDGNN_E_CUDA;
    const int wpb = 8;
    long long want = (rows + wpb - 1) / wpb;
    const int grid_x = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);      // 8 CTAs of 8 warps per SM, grid-stride
    cudaStream_t st = (cudaStream_t)stream;
    const int kmax = (pitch + 31) / 32;
#define LAUNCH(T, K) compact_from_raw_kernel<T, K><<<grid_x, 256, 0, st>>>((const T*)raw, (int)rows, n, grid, pitch, diag, err)
    if (raw_is_f64) {
        if (kmax <= 4) LAUNCH(double, 4); else if (kmax <= 8) LAUNCH(double, 8); else if (kmax <= 12) LAUNCH(double, 12); else LAUNCH(double, 17);
    } else {
        if (kmax <= 4) LAUNCH(float, 4); else if (kmax <= 8) LAUNCH(float, 8); else if (kmax <= 12) LAUNCH(float, 12); else LAUNCH(float, 17);
    }
#undef LAUNCH
    return cudaGetLastError() == cudaSuccess ? HDGNN_OK : HDGNN_E_CUDA;
}

extern "C" int hdgnn_eval_counts(int B, int Nc, const float* probs, const uint8_t* Y, int y_pitch, int64_t* counts, int64_t* auc,
                                 int auc_first, void* stream) {
    if (B < 1 || Nc < 2 || Nc > HDGNN_MAX_N || !probs || !Y || !counts || y_pitch < Nc) return HDGNN_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(counts, 0, (size_t)B * 8 * sizeof(int64_t), st) != cudaSuccess) return HDGNN_E_CUDA;
    if (auc && cudaMemsetAsync(auc, 0, (size_t)B * 2 * sizeof(int64_t), st) != cudaSuccess) return HDGNN_E_CUDA;
    const int Ncr = Nc * (Nc - 1);
    dim3 g((Ncr + EV_CHUNK - 1) / EV_CHUNK, B);
    eval_counts_kernel<<<g, EV_T, 0, st>>>(Nc, probs, Y, y_pitch, (unsigned long long*)counts, (unsigned long long*)auc,
                                          auc_first < 0 ? 0 : auc_first);
    return cudaGetLastError() == cudaSuccess ? HDGNN_OK : HDGNN_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Measured fp32 CUDA-core peaks for bench.py's roofline denominators (not part of the hot path): a register-resident chain
// of independent fused multiply-adds, scalar (FFMA) or packed (fma.rn.f32x2, SASS FFMA2 -- the form the pair sweeps use).
namespace hdgnn {
template <bool PACKED>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float* sink) {
    constexpr int NA = 8;
    if (PACKED) {
        unsigned long long acc[NA], m, c;
        asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(1.0f + seed * 1e-7f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(seed * 1e-3f));
#pragma unroll
        for (int i = 0; i < NA; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"((float)(threadIdx.x + i)), "f"((float)i));
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < NA; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(m), "l"(c));
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NA; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
        if (s == 12345.678f) sink[0] = s;
    } else {
        float acc[NA];
        const float m = 1.0f + seed * 1e-7f, c = seed * 1e-3f;
#pragma unroll
        for (int i = 0; i < NA; ++i) acc[i] = (float)(threadIdx.x + i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < NA; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(m), "f"(c));
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NA; ++i) s += acc[i];
        if (s == 12345.678f) sink[0] = s;
    }
}
}  // namespace hdgnn

extern "C" int hdgnn_measure_fp32_peak(int packed, float* tflops_out) {
    if (!tflops_out) return HDGNN_E_INVALID;
    int dev = 0, nsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return HDGNN_E_CUDA;
    float* sink = nullptr;
    if (cudaMalloc(&sink, 16) != cudaSuccess) return HDGNN_E_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = nsm * 8, block = 256;
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (packed) hdgnn::fma_peak_kernel<true><<<grid, block>>>(iters, 1.f, sink);
        else hdgnn::fma_peak_kernel<false><<<grid, block>>>(iters, 1.f, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return HDGNN_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)grid * block * iters * 8.0 * 2.0 * (packed ? 2.0 : 1.0);
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops_out = best;
    return HDGNN_OK;
}
