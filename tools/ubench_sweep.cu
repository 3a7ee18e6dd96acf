// Micro-benchmark of candidate inner loops for the entity forward sweep (N=200, RT=50, B=100).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_sweep tools/ubench_sweep.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

constexpr int HD = 20;
constexpr float NEG_BIG = -1e30f;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- variant B: warp = 4 channels x CW column segments, in-lane row accumulation ---------------
// 4 values -> after the call lanes with (lane & 7) == 0 hold the total of index ((lane>>4)&1) + 2*((lane>>3)&1)
__device__ __forceinline__ float reduce4(float a0, float a1, float a2, float a3, int lane) {
    const bool h16 = lane & 16;
    float k0 = h16 ? a1 : a0, s0 = h16 ? a0 : a1;
    float k1 = h16 ? a3 : a2, s1 = h16 ? a2 : a3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool h8 = lane & 8;
    float k = h8 ? k1 : k0, s = h8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

struct Args { const uint8_t* lab; int pitch, N, RT, S; const float* x; const float* par; float* RS; float* CSp; int reps; };

template <int CW, int NRG, int MODE>   // MODE 0 scalar, 1 f32x2 packed, 2 scalar + relu on the FMA pipe (t+|t|)
__global__ void __launch_bounds__(5 * NRG * 32) sweepB(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, RT = a.RT, pitch = a.pitch;
    const int b = blockIdx.y, s = blockIdx.x, r0 = s * RT, nrows = min(RT, N - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kg = warp % 5, rg = warp / 5, k0 = kg * 4;
    uint8_t* stage = smem;                                          // [RT][pitch]
    float* P01 = reinterpret_cast<float*>(smem + ((RT * pitch + 127) & ~127));   // [RT][2][20]
    float* rowacc = P01 + 2 * RT * HD;                              // [RT][20]
    float* cpart = rowacc + RT * HD;                                // [NRG][CW*32][20]
    const float* par = a.par;     // u[20] v[20] b[20] w0[20] w1[20]
    const float* xb = a.x + (size_t)b * N;
    for (int i = tid; i < nrows * pitch; i += blockDim.x) stage[i] = a.lab[((size_t)b * N + r0) * pitch + i];
    for (int idx = tid; idx < nrows * HD; idx += blockDim.x) {
        const int r = idx / HD, k = idx - r * HD;
        const float p0 = fmaf(xb[r0 + r], par[k], par[40 + k] + par[60 + k]);
        P01[(r * 2) * HD + k] = p0;
        P01[(r * 2 + 1) * HD + k] = p0 + (par[80 + k] - par[60 + k]);
    }
    float Q[CW][4], col[CW][4];
#pragma unroll
    for (int sg = 0; sg < CW; ++sg) {
        const int j = sg * 32 + lane;
        const float xj = j < N ? xb[j] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { Q[sg][k] = j < N ? xj * par[20 + k0 + k] : NEG_BIG; col[sg][k] = 0.f; }
    }
    __syncthreads();
    for (int rep = 0; rep < a.reps; ++rep)
    for (int r = rg; r < nrows; r += NRG) {
        float rp[4] = {0.f, 0.f, 0.f, 0.f};
        const uint8_t* lrow = stage + r * pitch + lane;
        const float* prow = P01 + (r * 2) * HD + k0;
#pragma unroll
        for (int sg = 0; sg < CW; ++sg) {
            const int j = sg * 32 + lane;
            const bool lab = (j < N) && (j != r0 + r) && lrow[sg * 32] != 0;
            const float4 p = *reinterpret_cast<const float4*>(prow + (lab ? HD : 0));
            const float P[4] = {p.x, p.y, p.z, p.w};
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                    float t0, t1;
                    asm("{ .reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; add.f32x2 c, a, b; mov.b64 {%0, %1}, c; }"
                        : "=f"(t0), "=f"(t1) : "f"(P[k]), "f"(P[k + 1]), "f"(Q[sg][k]), "f"(Q[sg][k + 1]));
                    const float h0 = fmaxf(t0, 0.f), h1 = fmaxf(t1, 0.f);
                    asm("{ .reg .b64 a, b, c; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; add.f32x2 c, a, b; mov.b64 {%0, %1}, c; }"
                        : "+f"(col[sg][k]), "+f"(col[sg][k + 1]) : "f"(h0), "f"(h1));
                    asm("{ .reg .b64 a, b, c; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; add.f32x2 c, a, b; mov.b64 {%0, %1}, c; }"
                        : "+f"(rp[k]), "+f"(rp[k + 1]) : "f"(h0), "f"(h1));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float t = P[k] + Q[sg][k];
                    const float h = MODE == 2 ? t + fabsf(t) : fmaxf(t, 0.f);
                    col[sg][k] += h;
                    rp[k] += h;
                }
            }
        }
        const float tot = reduce4(rp[0], rp[1], rp[2], rp[3], lane);
        if ((lane & 7) == 0) rowacc[r * HD + k0 + ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1)] = tot;
    }
    // column partials per row group -> smem -> global
#pragma unroll
    for (int sg = 0; sg < CW; ++sg)
#pragma unroll
        for (int k = 0; k < 4; ++k) cpart[((size_t)rg * CW * 32 + sg * 32 + lane) * HD + k0 + k] = col[sg][k];
    __syncthreads();
    const float sc = MODE == 2 ? 0.5f : 1.f;
    for (int idx = tid; idx < N * HD; idx += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < NRG; ++g) v += cpart[(size_t)g * CW * 32 * HD + idx];
        a.CSp[((size_t)b * a.S + s) * N * HD + idx] = v * sc;
    }
    for (int idx = tid; idx < nrows * HD; idx += blockDim.x) a.RS[((size_t)b * N + r0) * HD + idx] = rowacc[idx] * sc;
}

int main() {
    const int B = 100, N = 200, pitch = 208, S = 4, RT = 50;
    std::vector<uint8_t> lab((size_t)B * N * pitch);
    std::vector<float> x((size_t)B * N), par(100);
    srand(1);
    for (auto& v : lab) v = (rand() % 100) < 5;
    for (auto& v : x) v = (float)(rand() % 10);
    for (auto& v : par) v = 0.2f * ((rand() % 2001) / 1000.f - 1.f);
    uint8_t* dl; float *dx, *dp, *drs, *dcs;
    CK(cudaMalloc(&dl, lab.size())); CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dp, 400));
    CK(cudaMalloc(&drs, (size_t)B * N * HD * 4)); CK(cudaMalloc(&dcs, (size_t)B * S * N * HD * 4));
    CK(cudaMemcpy(dl, lab.data(), lab.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, par.data(), 400, cudaMemcpyHostToDevice));
    Args a{dl, pitch, N, RT, S, dx, dp, drs, dcs, 1};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ref;
    auto run = [&](const char* name, auto kern, int threads, size_t smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float tm[2];
        for (int pass = 0; pass < 2; ++pass) {
            a.reps = pass == 0 ? 5 : 1;
            for (int i = 0; i < 5; ++i) kern<<<dim3(S, B), threads, smem>>>(a);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int i = 0; i < 50; ++i) kern<<<dim3(S, B), threads, smem>>>(a);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&tm[pass], e0, e1);
        }
        const float ms = tm[1];
        printf("   marginal sweep %.2f us, fixed %.2f us | ", (tm[0] - tm[1]) * 1000 / 50 / 4, (tm[1] - (tm[0] - tm[1]) / 4) * 1000 / 50);
        std::vector<float> rs((size_t)B * N * HD), cs((size_t)B * S * N * HD);
        CK(cudaMemcpy(rs.data(), drs, rs.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cs.data(), dcs, cs.size() * 4, cudaMemcpyDeviceToHost));
        double sum = 0, sc = 0; for (float v : rs) sum += v; for (float v : cs) sc += v;
        double maxd = 0;
        if (ref.empty()) ref = rs; else for (size_t i = 0; i < rs.size(); ++i) { double d = fabs(rs[i] - ref[i]) / (fabs(ref[i]) + 1e-3); if (d > maxd) maxd = d; }
        printf("%-28s %7.2f us/launch   RS sum %.6e  CS sum %.6e  max rel diff vs first %.2e\n", name, ms * 1000 / 50, sum, sc, maxd);
    };
    auto smemB = [&](int nrg, int cw) { return (size_t)((RT * pitch + 127) & ~127) + (size_t)RT * HD * 4 * 3 + (size_t)nrg * cw * 32 * HD * 4; };
    run("B scalar  NRG=2 (10 warps)", sweepB<7, 2, 0>, 320, smemB(2, 7));
    run("B scalar  NRG=3 (15 warps)", sweepB<7, 3, 0>, 480, smemB(3, 7));
    run("B scalar  NRG=4 (20 warps)", sweepB<7, 4, 0>, 640, smemB(4, 7));
    run("B f32x2   NRG=2", sweepB<7, 2, 1>, 320, smemB(2, 7));
    run("B f32x2   NRG=3", sweepB<7, 3, 1>, 480, smemB(3, 7));
    run("B f32x2   NRG=4", sweepB<7, 4, 1>, 640, smemB(4, 7));
    run("B t+|t|   NRG=3", sweepB<7, 3, 2>, 480, smemB(3, 7));
    run("B t+|t|   NRG=4", sweepB<7, 4, 2>, 640, smemB(4, 7));
    return 0;
}
