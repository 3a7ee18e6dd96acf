// Micro-benchmark of the general-path pooling-backward column gather (mid2.cuh phase K), one CTA of 640 threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
template <int V>
__global__ void __launch_bounds__(640, 1) gk(const float* dlg, float* out, long long* clk, int Ne, int Lb) {
    extern __shared__ float dl[];
    for (int i = threadIdx.x; i < 4 * Ne; i += blockDim.x) { dl[i] = dlg[i]; if ((i & 3) == 1) dl[4 * Ne + (i >> 2)] = dlg[i]; }
    __syncthreads();
    const long long t0 = clock64();
    const int tid = threadIdx.x, nm1 = Ne - 1;
    const int n = nm1, m = Lb - 1, qmax = Lb * m, d = n - m;
    int nchunk = 640 / Ne;
    nchunk = nchunk < 1 ? 1 : (nchunk > 4 ? 4 : nchunk);
    for (int t = tid; t < nchunk * Ne; t += 640) {
        const int c = t / Ne, me = t - c * Ne;
        const int lo = V == 0 ? (int)(((long long)c * Ne) / nchunk) : (c * Ne) / nchunk;
        const int hi = V == 0 ? (int)(((long long)(c + 1) * Ne) / nchunk) : ((c + 1) * Ne) / nchunk;
        float acc = 0.f;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int g0 = part == 0 ? lo : max(lo, me + 1), g1 = part == 0 ? min(hi, me) : hi;
            if (g0 >= g1) continue;
            int q = g0 * n + me - (part == 0 ? 1 : 0);
            int li = q / m, sloc = q - li * m;
            if (V == 3) {
                // counted, branch-free (d < m: at most one wrap per step), channel-planar dl (dl1 = dl + 4 Ne)
                int trips = g1 - g0;
                if (q >= qmax) trips = 0; else { const int lim = (qmax - 1 - q) / n + 1; trips = trips < lim ? trips : lim; }
                const float* dl1 = dl + 4 * Ne;
#pragma unroll 4
                for (int it = 0; it < trips; ++it) {
                    const int lj = sloc + (sloc >= li);
                    acc += dl1[li] + dl1[lj];
                    sloc += d;
                    const int w = sloc >= m;
                    sloc -= w ? m : 0;
                    li += 1 + w;
                }
            } else if (V == 2) {
                // trip count known up front: no data-dependent exit, unrollable
                int trips = g1 - g0;
                if (q >= qmax) trips = 0; else { const int lim = (qmax - 1 - q) / n + 1; trips = trips < lim ? trips : lim; }
#pragma unroll 4
                for (int it = 0; it < trips; ++it) {
                    const int lj = sloc + (sloc >= li);
                    acc += dl[4 * li + 1] + dl[4 * lj + 1];
                    sloc += d; ++li;
                    while (sloc >= m) { sloc -= m; ++li; }
                }
            } else {
                for (int gi = g0; gi < g1 && q < qmax; ++gi) {
                    const int lj = sloc + (sloc >= li);
                    acc += dl[4 * li + 1] + dl[4 * lj + 1];
                    q += n; sloc += d; ++li;
                    while (sloc >= m) { sloc -= m; ++li; }
                }
            }
        }
        out[t] = acc;
    }
    __syncthreads();
    if (tid == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
    const int Ne = 200;
    std::vector<float> h(4 * Ne);
    for (int i = 0; i < 4 * Ne; ++i) h[i] = (float)(i % 7) * 0.25f;
    float *d, *o; long long* c;
    cudaMalloc(&d, 4 * Ne * 4); cudaMalloc(&o, 4096 * 4); cudaMalloc(&c, 8 * 148);
    cudaMemcpy(d, h.data(), 4 * Ne * 4, cudaMemcpyHostToDevice);
    for (int L : {199, 175, 150, 125, 100}) {
        long long r[4];
        double s[4];
        for (int v = 0; v < 4; ++v) {
            for (int rep = 0; rep < 3; ++rep) {
                if (v == 0) gk<0><<<100, 640, 5 * Ne * 4>>>(d, o, c, Ne, L);
                if (v == 1) gk<1><<<100, 640, 5 * Ne * 4>>>(d, o, c, Ne, L);
                if (v == 2) gk<2><<<100, 640, 5 * Ne * 4>>>(d, o, c, Ne, L);
                if (v == 3) gk<3><<<100, 640, 5 * Ne * 4>>>(d, o, c, Ne, L);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&r[v], c, 8, cudaMemcpyDeviceToHost);
            std::vector<float> ho(600);
            cudaMemcpy(ho.data(), o, 600 * 4, cudaMemcpyDeviceToHost);
            s[v] = 0; for (float x : ho) s[v] += x;
        }
        printf("L=%d cycles: v0 %lld  v1(no 64-bit div) %lld  v2(counted loop) %lld  v3(branch-free, planar) %lld   sums %.3f %.3f %.3f %.3f  %s\n", L, r[0], r[1], r[2], r[3], s[0], s[1], s[2], s[3],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
