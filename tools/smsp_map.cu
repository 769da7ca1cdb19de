// Which SM sub-partition does warp w of a CTA land on?  512 CTAs x 128 threads with 54 KB of dynamic shared memory (the
// C4 launch shape): every warp records %smid, %warpid; then a contention test: ONE busy warp per CTA (issue-bound loop),
// chosen as warp 0 / warp (slot & 3) where slot = arrival order of the CTA on its SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
__global__ void __launch_bounds__(128, 4) map_kernel(int* out) {
    extern __shared__ double sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2] = smid;
        out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2 + 1] = wid;
    }
    // keep the CTA resident for a while so that co-residency is the steady state
    long long t0 = clock64();
    while (clock64() - t0 < 200000) {}
    if (sm[threadIdx.x] == 12345.0) out[0] = 1;
}
__global__ void __launch_bounds__(128, 4) busy_kernel(int mode, int* cnt, double* out, int iters) {
    extern __shared__ double sm[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    __shared__ int slot;
    if (threadIdx.x == 0) slot = atomicAdd(&cnt[smid], 1);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int busy = mode == 0 ? 0 : (mode == 1 ? (slot & 3) : (blockIdx.x & 3));
    if (warp != busy) return;
    // issue-heavy: 4 independent integer/fp chains (ILP 4) -> one warp wants ~1 issue / cycle
    double a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    for (int i = 0; i < iters; ++i) {
        a = fma(a, 1.0000001, 0.5); b = fma(b, 1.0000001, 0.5); c = fma(c, 1.0000001, 0.5); d = fma(d, 1.0000001, 0.5);
    }
    if (a + b + c + d == 1.2345) out[0] = a;
}
int main() {
    int* d; cudaMalloc(&d, 512 * 4 * 2 * sizeof(int));
    cudaFuncSetAttribute(map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 54048);
    cudaFuncSetAttribute(busy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 54048);
    map_kernel<<<512, 128, 54048>>>(d);
    std::vector<int> h(512 * 8);
    cudaMemcpy(h.data(), d, h.size() * sizeof(int), cudaMemcpyDeviceToHost);
    // per SM: list (block, warp -> warpid)
    for (int sm = 0; sm < 3; ++sm) {
        printf("SM %d:", sm);
        for (int b = 0; b < 512; ++b)
            if (h[b * 8] == sm) { printf("  cta %d warpids", b); for (int w = 0; w < 4; ++w) printf(" %d", h[(b * 4 + w) * 2 + 1]); }
        printf("\n");
    }
    int hist[4][4] = {};
    for (int b = 0; b < 512; ++b) for (int w = 0; w < 4; ++w) hist[w][h[(b * 4 + w) * 2 + 1] & 3]++;
    for (int w = 0; w < 4; ++w) printf("warp %d -> (warpid & 3) histogram: %d %d %d %d\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
    int percount[8] = {};
    { std::vector<int> c(200, 0); for (int b = 0; b < 512; ++b) c[h[b * 8]]++; for (int s = 0; s < 200; ++s) percount[std::min(c[s], 7)]++; }
    printf("CTAs per SM histogram (0..7):"); for (int i = 0; i < 8; ++i) printf(" %d", percount[i]); printf("\n");
    int* cnt; cudaMalloc(&cnt, 256 * sizeof(int));
    double* o; cudaMalloc(&o, 8);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(cnt, 0, 256 * sizeof(int));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            busy_kernel<<<512, 128, 54048>>>(mode, cnt, o, 200000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("busy warp mode %d (0: warp 0, 1: slot&3, 2: blockIdx&3): %.3f ms\n", mode, ms);
        }
    }
    return 0;
}
