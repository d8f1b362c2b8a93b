// probe: can two CUDA green contexts give two kernels disjoint SM sets on this box, with plain runtime-API launches on their
// streams and cudaMalloc'ed memory of the primary context? Prints the SM ids each kernel ran on and whether they overlapped
// in time. Driver entry points come from cudaGetDriverEntryPoint: nothing links libcuda.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#define DRV(name) \
    decltype(&name) p_##name = nullptr; \
    { cudaDriverEntryPointQueryResult qr; void* f = nullptr; \
      if (cudaGetDriverEntryPoint(#name, &f, cudaEnableDefault, &qr) != cudaSuccess || !f) { printf("no entry point %s\n", #name); return 2; } \
      p_##name = reinterpret_cast<decltype(&name)>(f); }
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("%s -> %d\n", #x, (int)r_); return 3; } } while (0)

__global__ void spin(unsigned* smids, long long* t, long long cycles) {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        smids[blockIdx.x] = s;
        long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        t[2 * blockIdx.x] = g;
    }
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 0) {
        long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        t[2 * blockIdx.x + 1] = g;
    }
}

int main(int argc, char** argv) {
    const unsigned want = argc > 1 ? (unsigned)atoi(argv[1]) : 32;
    cudaFree(0);
    DRV(cuDeviceGet) DRV(cuDeviceGetDevResource) DRV(cuDevSmResourceSplitByCount) DRV(cuDevResourceGenerateDesc)
    DRV(cuGreenCtxCreate) DRV(cuGreenCtxStreamCreate) DRV(cuGreenCtxDestroy)
    CUdevice dev;
    CK(p_cuDeviceGet(&dev, 0));
    CUdevResource all, grp[4], rem;
    CK(p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", all.sm.smCount);
    unsigned n = 1;
    CK(p_cuDevSmResourceSplitByCount(grp, &n, &all, &rem, 0, want));
    printf("split minCount %u -> %u group(s) of %u SMs, remaining %u SMs\n", want, n, grp[0].sm.smCount, rem.sm.smCount);
    CUdevResourceDesc dA, dB;
    CK(p_cuDevResourceGenerateDesc(&dA, &grp[0], 1));
    CK(p_cuDevResourceGenerateDesc(&dB, &rem, 1));
    CUgreenCtx gA, gB;
    CK(p_cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CK(p_cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream sA, sB;
    CK(p_cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
    CK(p_cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
    const int nb = 600;
    unsigned *smA, *smB; long long *tA, *tB;
    cudaMalloc(&smA, nb * 4); cudaMalloc(&smB, nb * 4); cudaMalloc(&tA, nb * 16); cudaMalloc(&tB, nb * 16);
    // kernel with a big dynamic shared-memory footprint in B (one CTA per SM), small CTAs in A
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    spin<<<nb, 512, 200 * 1024, (cudaStream_t)sB>>>(smB, tB, 2000000);
    spin<<<nb, 32, 0, (cudaStream_t)sA>>>(smA, tA, 2000000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    std::vector<unsigned> hA(nb), hB(nb); std::vector<long long> hta(2 * nb), htb(2 * nb);
    cudaMemcpy(hA.data(), smA, nb * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hB.data(), smB, nb * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hta.data(), tA, nb * 16, cudaMemcpyDeviceToHost); cudaMemcpy(htb.data(), tB, nb * 16, cudaMemcpyDeviceToHost);
    std::set<unsigned> SA(hA.begin(), hA.end()), SB(hB.begin(), hB.end());
    int common = 0;
    for (unsigned s : SA) common += SB.count(s);
    long long a0 = hta[0], a1 = hta[1], b0 = htb[0], b1 = htb[1];
    for (int i = 0; i < nb; i++) { a0 = std::min(a0, hta[2 * i]); a1 = std::max(a1, hta[2 * i + 1]); b0 = std::min(b0, htb[2 * i]); b1 = std::max(b1, htb[2 * i + 1]); }
    printf("A ran on %zu SMs, B on %zu SMs, %d in common\n", SA.size(), SB.size(), common);
    printf("A: %.3f ms .. %.3f ms   B: %.3f .. %.3f ms (relative to the earlier start)\n", (a0 - std::min(a0, b0)) / 1e6, (a1 - std::min(a0, b0)) / 1e6,
           (b0 - std::min(a0, b0)) / 1e6, (b1 - std::min(a0, b0)) / 1e6);
    p_cuGreenCtxDestroy(gA); p_cuGreenCtxDestroy(gB);
    return 0;
}
