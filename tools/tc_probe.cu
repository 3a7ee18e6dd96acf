// Probe of the hand-written tcgen05 path used by k_conv.cu: D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T with bf16
// operands in shared memory in the canonical no-swizzle K-major layout, one CTA.  Checks the descriptor encodings
// against a CPU product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { for (uint32_t s = 0; !mbar_try(bar, parity); ++s) if (s > (1u << 24)) __trap(); }

// shared-memory matrix descriptor, no swizzle, K-major: 8 x 16-byte core matrices; LBO = byte step to the next 8 K elements,
// SBO = byte step to the next 8 rows (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, shape M x N (UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int N, int K>
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    // A: [128][K] row-major, B: [N][K] row-major (global) -> smem canonical: [K/8][rows][8] bf16
    extern __shared__ __align__(128) unsigned char smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sB = sA + 128 * K;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * K);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < 128 * K; idx += 128) { const int r = idx / K, k = idx - r * K; sA[((k >> 3) * 128 + r) * 8 + (k & 7)] = A[idx]; }
    for (int idx = tid; idx < N * K; idx += 128) { const int r = idx / K, k = idx - r * K; sB[((k >> 3) * N + r) * 8 + (k & 7)] = B[idx]; }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (rows), 16 columns at a time
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N) : "memory");
}

int main() {
    constexpr int N = 64, K = 208;
    std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K);
    srand(3);
    for (int i = 0; i < 128 * K; ++i) { fA[i] = (rand() % 100) < 20 ? 1.f : 0.f; hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001) / 1000.f - 1.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * N * 4);
    const size_t smem = (128 * K + N * K) * 2 + 64;
    cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<N, K><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - hD[m * N + n])); maxref = fmax(maxref, fabs(ref));
        }
    printf("tcgen05 probe: max abs err %.3e (max |ref| %.3f) -> %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
    printf("D[0][0..3] = %f %f %f %f\n", hD[0], hD[1], hD[2], hD[3]);
    return 0;
}
