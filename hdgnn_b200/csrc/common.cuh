// Shared device helpers for the HD-GNN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdgnn {

constexpr int HD = 20;            // h_size / De_e / De_er (model_2.py:163, main.py:31-32)
constexpr float NEG_BIG = -1e30f; // additive mask: relu(x + NEG_BIG) == 0

// Offsets (in floats) into the flat parameter blob; -1 = block absent in this variant.
// Order = TF variable-creation order of build_model (model_2.py:86-121, model_4.py:84-113).
struct ParamOff {
    int ent_w1, ent_b1, ent_w5, ent_b5, nod_w1, nod_b1, nod_w2, nod_b2;
    int edg_w11, edg_w12, edg_b1, edg_w2, edg_b2, eup_w1, eup_b1, eup_w2, eup_b2;
    int hnk_w1, hnk_b1, hnk_w2, hnk_b2, scr_w1, scr_b1, scr_w2, scr_b2;
    int theta1, theta2, total;
};

// The "pair layer + head" weights shared by the hunk stage (hnk_w2/hnk_b2 + scr_*) and the
// entity-edge branch (edg_w2/edg_b2 + eup_*).
struct HeadOff {
    int w2p, b2p;   // second (linear) layer of the pair MLP: (20,20), (20)
    int w1h, b1h;   // first layer of the head: (22,20), (20); rows 0,1 = label one-hot
    int w2h, b2h;   // (20,2), (2)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk (TMA) copy, sm_90+ PTX --------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// global -> shared::cta bulk copy; dst, src 16-byte aligned, bytes a multiple of 16 (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost bulk copy traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}

// ---- programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute
// may start while its predecessor drains; it must not touch the predecessor's outputs before pdl_wait().
// Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- thread-block clusters: rank, barrier, distributed shared memory (sm_90+) ---------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster; release / acquire: shared-memory writes before it are visible to the peers' reads after it
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t peer_smem(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_peer4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// ---- reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of 4 per-lane values per row over the 32 lanes of a warp ("transpose reduce"):
// after the call every lane holds the total of row ((lane>>4)&1)*2 + ((lane>>3)&1).
// Fixed exchange order => bitwise deterministic.
__device__ __forceinline__ float warp_rowsum4(float r0, float r1, float r2, float r3, int lane) {
    const bool hi16 = lane & 16;
    float k0 = hi16 ? r2 : r0, k1 = hi16 ? r3 : r1;
    float s0 = hi16 ? r0 : r2, s1 = hi16 ? r1 : r3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool hi8 = lane & 8;
    float k = hi8 ? k1 : k0, s = hi8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// Block-wide sum in a fixed order (warp tree, then warps in index order). `scratch` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += scratch[w];
    return t;
}

__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ void copy_to_smem(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// q = i*(n-1) + (j - [j>i])  ->  (i, j); q < 2^18 so the float reciprocal is exact after fix-up
__device__ __forceinline__ void unflat_pair(int q, int nm1, float inv, int& i, int& j) {
    int gi = (int)(((float)q + 0.5f) * inv);
    int gr = q - gi * nm1;
    if (gr < 0) { --gi; gr += nm1; }
    else if (gr >= nm1) { ++gi; gr -= nm1; }
    i = gi;
    j = gr + (gr >= gi);
}


}  // namespace hdgnn
