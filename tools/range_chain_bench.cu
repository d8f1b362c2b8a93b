// microbenchmark: cycles per record of the serial range recurrence, one warp, several formulations
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void step(uint32_t qb, uint32_t c, uint32_t& range, uint32_t& cnt) {
    if (V == 0) {           // compare + select
        const uint32_t sp = qb + 1;
        const uint32_t x = range * sp + c;
        const bool sh = x < 0x10000u;
        range = sh ? (x & 0xFFFF00u) : (x >> 8);
        cnt += sh;
    } else if (V == 1) {    // biased range rb = range - 0x100, two IMADs + min
        // range here holds rb. x = (rb + 0x100) * sp + c
        const uint32_t sp = qb + 1;
        const uint32_t k1 = (sp << 8) + c - 0x10000u;
        const uint32_t y = range * sp + k1;             // x - 0x10000
        const uint32_t y2 = range * sp + (k1 - 0x100u); // x - 0x10100
        const uint32_t a = y >> 8;                      // (x >> 8) - 0x100, huge if renorm needed
        const uint32_t b = y2 & 0xFF00u;                // (x & 0xFF00) - 0x100 in the renorm case
        cnt += (b < a) ? ((a >> 16) ? 1u : 0u) : 0u;
        range = min(a, b);
    } else if (V == 2) {    // V1 with the count taken from the sign of y
        const uint32_t sp = qb + 1;
        const uint32_t k1 = (sp << 8) + c - 0x10000u;
        const uint32_t y = range * sp + k1;
        const uint32_t y2 = range * sp + (k1 - 0x100u);
        const uint32_t a = y >> 8;
        const uint32_t b = y2 & 0xFF00u;
        cnt += y >> 31;
        range = min(a, b);
    } else if (V == 3) {    // shift amount from the high half
        const uint32_t sp = qb + 1;
        const uint32_t x = range * sp + c;
        const uint32_t hi = x >> 16;
        const uint32_t s = hi ? 8u : 0u;
        range = (x >> s) & 0xFFFF00u ? ((x >> s) & (hi ? 0xFFFFu : 0xFF00u)) : 0x100u;
        cnt += hi ? 0u : 1u;
    }
}

template <int V, int ILP>
__global__ void bench(const uint32_t* __restrict__ q, const uint32_t* __restrict__ bits, int nblk, uint32_t* out, long long* cyc) {
    __shared__ uint32_t sq[1024];     // 4096 records
    __shared__ uint32_t sb[128];
    for (int i = threadIdx.x; i < 1024; i += 32) sq[i] = q[i];
    for (int i = threadIdx.x; i < 128; i += 32) sb[i] = bits[i];
    __syncwarp();
    uint32_t range[ILP], cnt[ILP];
    for (int j = 0; j < ILP; j++) { range[j] = V == 1 || V == 2 ? 0xFF00u - 0x100u : 0xFF00u; cnt[j] = 0; }
    const long long t0 = clock64();
    for (int b = 0; b < nblk; b++) {
        const int base = ((b * 8 + threadIdx.x) & 31) * 32;    // 32 words = 128 records
        const uint32_t* bw = sb + (base >> 3);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint4 w = *reinterpret_cast<const uint4*>(sq + base + u * 4);
            const uint32_t nb = ~(bw[u >> 1] >> ((u & 1) * 16));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t qb = __byte_perm(ww[k >> 2], 0, 0x4440 | (k & 3));
                const uint32_t t = nb << (7 - (k & 7));
                const uint32_t c = __byte_perm(t, 0, 0x4440 | 8 | (k >> 3));
#pragma unroll
                for (int j = 0; j < ILP; j++) step<V>(qb ^ (j * 3), c, range[j], cnt[j]);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
    for (int j = 0; j < ILP; j++) r += range[j] + cnt[j];
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// V4: 16-bit records (q, nb) in byte pairs, x + 0xFF000000 = dp2a(A, w, C) with A = range | 0x00FF0000, C = A + 0xFE010000:
// no extraction instructions, the renormalisation as two predicated single-instruction updates of A
template <int ILP, int NW>
__global__ void bench_dp2a(const uint32_t* __restrict__ q, const uint32_t* __restrict__ bits, int nblk, uint32_t* out, long long* cyc) {
    __shared__ uint32_t sq[1][1024];
    __shared__ uint32_t sb[1][128];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int i = lane; i < 1024; i += 32) sq[0][i] = q[i];
    for (int i = lane; i < 128; i += 32) sb[0][i] = bits[i];
    __syncthreads();
    uint32_t A[ILP], C[ILP], cnt[ILP];
    const uint32_t K = 0xFE010000u, H = 0x00FF0000u, M = 0x0000FF00u, T = 0xFF010000u;
    for (int j = 0; j < ILP; j++) { A[j] = 0xFF00u | H; C[j] = A[j] + K; cnt[j] = 0; }
    const long long t0 = clock64();
    for (int b = 0; b < nblk; b++) {
        const int base = ((b * 8 + lane + wp) & 31) * 32;
        const uint32_t* bw = sb[0] + (base >> 3);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint4 w = *reinterpret_cast<const uint4*>(sq[0] + base + u * 4);
            const uint32_t nb = ~(bw[u >> 1] >> ((u & 1) * 16));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            // expand: 16 q bytes + 16 bits -> 8 words of (q, nb, q, nb)
            uint32_t P[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // bits 4k..4k+3 of nb -> bytes 0/1
                const uint32_t n4 = (nb >> (4 * k)) & 15u;
                const uint32_t nby = (n4 * 0x00204081u) & 0x01010101u;
                P[2 * k] = __byte_perm(ww[k], nby, 0x5140);
                P[2 * k + 1] = __byte_perm(ww[k], nby, 0x7362);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
#pragma unroll
                for (int j = 0; j < ILP; j++) {
                    const uint32_t wv = P[k >> 1] ^ (uint32_t)(j * 3);
                    const uint32_t x = (k & 1) ? __dp2a_hi(A[j], wv, C[j]) : __dp2a_lo(A[j], wv, C[j]);
                    const bool sh = x < T;
                    A[j] = sh ? ((x & M) | H) : (x >> 8);
                    C[j] = A[j] + K;
                    cnt[j] += sh;
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
    for (int j = 0; j < ILP; j++) r += (A[j] & 0xFFFFu) + cnt[j];
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// several warps of V0 on one SM (NW warps, one CTA): what sharing a scheduler costs the chain
template <int NW>
__global__ void bench_v0_multi(const uint32_t* __restrict__ q, const uint32_t* __restrict__ bits, int nblk, uint32_t* out, long long* cyc) {
    __shared__ uint32_t sq[1][1024];
    __shared__ uint32_t sb[1][128];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int i = lane; i < 1024; i += 32) sq[0][i] = q[i];
    for (int i = lane; i < 128; i += 32) sb[0][i] = bits[i];
    __syncthreads();
    uint32_t range = 0xFF00u, cnt = 0;
    const long long t0 = clock64();
    for (int b = 0; b < nblk; b++) {
        const int base = ((b * 8 + lane + wp) & 31) * 32;
        const uint32_t* bw = sb[0] + (base >> 3);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint4 w = *reinterpret_cast<const uint4*>(sq[0] + base + u * 4);
            const uint32_t nb = ~(bw[u >> 1] >> ((u & 1) * 16));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t qb = __byte_perm(ww[k >> 2], 0, 0x4440 | (k & 3));
                const uint32_t t = nb << (7 - (k & 7));
                const uint32_t c = __byte_perm(t, 0, 0x4440 | 8 | (k >> 3));
                step<0>(qb, c, range, cnt);
            }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = range + cnt;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int NW>
void run_dp2a(const uint32_t* q, const uint32_t* bits, uint32_t* out, long long* cyc) {
    const int nblk = 4096;
    bench_dp2a<ILP, NW><<<1, 32 * NW>>>(q, bits, nblk, out, cyc);
    bench_dp2a<ILP, NW><<<1, 32 * NW>>>(q, bits, nblk, out, cyc);
    cudaDeviceSynchronize();
    long long h; uint32_t ho[32];
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, out, 128, cudaMemcpyDeviceToHost);
    printf("%-28s ILP %d, %d warps/SM: %.2f cycles/record/chain-step  [chk %08x] %s\n", "V4 dp2a", ILP, NW, (double)h / (nblk * 128.0), ho[0], cudaGetErrorString(cudaGetLastError()));
}
template <int NW>
void run_v0_multi(const uint32_t* q, const uint32_t* bits, uint32_t* out, long long* cyc) {
    const int nblk = 4096;
    bench_v0_multi<NW><<<1, 32 * NW>>>(q, bits, nblk, out, cyc);
    bench_v0_multi<NW><<<1, 32 * NW>>>(q, bits, nblk, out, cyc);
    cudaDeviceSynchronize();
    long long h; uint32_t ho[32];
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, out, 128, cudaMemcpyDeviceToHost);
    printf("%-28s %d warps/SM: %.2f cycles/record  [chk %08x]\n", "V0 isetp/sel", NW, (double)h / (nblk * 128.0), ho[0]);
}

template <int V, int ILP>
void run(const uint32_t* q, const uint32_t* bits, uint32_t* out, long long* cyc, const char* name) {
    const int nblk = 4096;
    bench<V, ILP><<<1, 32>>>(q, bits, nblk, out, cyc);
    bench<V, ILP><<<1, 32>>>(q, bits, nblk, out, cyc);
    cudaDeviceSynchronize();
    long long h; uint32_t ho[32];
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, out, 128, cudaMemcpyDeviceToHost);
    printf("%-28s ILP %d: %.2f cycles/record/chain-step (%.2f per record)  [chk %08x]\n", name, ILP, (double)h / (nblk * 128.0), (double)h / (nblk * 128.0 * ILP), ho[0]);
}

int main() {
    uint32_t hq[1024], hb[128];
    uint32_t s = 12345;
    for (int i = 0; i < 1024; i++) { uint32_t w = 0; for (int k = 0; k < 4; k++) { s = s * 1664525u + 1013904223u; w |= ((s >> 24) % 255u) << (8 * k); } hq[i] = w; }
    for (int i = 0; i < 128; i++) { s = s * 1664525u + 1013904223u; hb[i] = s; }
    uint32_t *q, *bits, *out; long long* cyc;
    cudaMalloc(&q, 4096); cudaMalloc(&bits, 512); cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    cudaMemcpy(q, hq, 4096, cudaMemcpyHostToDevice); cudaMemcpy(bits, hb, 512, cudaMemcpyHostToDevice);
    run<0, 1>(q, bits, out, cyc, "V0 isetp/sel");
    run<0, 2>(q, bits, out, cyc, "V0 isetp/sel");
    run<1, 1>(q, bits, out, cyc, "V1 biased min");
    run<2, 1>(q, bits, out, cyc, "V2 biased min, sign count");
    run<2, 2>(q, bits, out, cyc, "V2 biased min, sign count");
    run<2, 3>(q, bits, out, cyc, "V2 biased min, sign count");
    run_v0_multi<1>(q, bits, out, cyc);
    run_v0_multi<4>(q, bits, out, cyc);
    run_v0_multi<8>(q, bits, out, cyc);
    run_v0_multi<16>(q, bits, out, cyc);
    run_dp2a<1, 1>(q, bits, out, cyc);
    run_dp2a<2, 1>(q, bits, out, cyc);
    run_dp2a<1, 4>(q, bits, out, cyc);
    run_dp2a<1, 8>(q, bits, out, cyc);
    run_dp2a<1, 16>(q, bits, out, cyc);
    return 0;
}
