// sm_100a FLAC encoder kernel (kernels K5 + K6 fused): one CTA per block (= one FLAC frame).
//
// Bitstream = exact inverse of the reference's vendored libFLAC 1.3.2 decoder:
//   frame header      /root/reference/Source/Lib/ThirdParty/flac/src/libFLAC/stream_decoder.c:2159-2466 (read_frame_header_)
//   subframes         stream_decoder.c:2468-2540 (read_subframe_), fixed predictor fixed.c:336-395
//   partitioned Rice  stream_decoder.c:2745-2788 (read_residual_partitioned_rice_), bitreader.c:744
//   CRC-8 / CRC-16    crc.c:366-376; checked at stream_decoder.c:2075-2127
// Per channel: constant test, fixed-predictor search (orders 0..4, criterion of fixed.c:217-273), zig-zag residual, exact
// search of the partition order 0..8 with Rice parameter floor(log2(mean)) per partition, VERBATIM fallback; then the bits:
// every thread knows the bit offset of its samples from a block-wide prefix sum and ORs its codes into the (zeroed) frame
// buffer held as big-endian 32-bit words. Frame CRC-16 in parallel (the CRC is linear: crc(A|B) = crc(A)*x^(8|B|) + crc(B)).
#include <cuda_runtime.h>

#include <cstdint>

#include "flac_kernels.cuh"

namespace b200 {

namespace {

constexpr int kT = kFlacThreads;

__device__ __forceinline__ void put_bits(uint32_t* W, uint64_t bitpos, uint32_t v, int n) {   // n in 1..32, MSB first
    const uint64_t w = bitpos >> 5;
    const int s = (int)(bitpos & 31);
    if (n < 32) v &= (1u << n) - 1u;
    if (s + n <= 32) atomicOr(&W[w], v << (32 - s - n));
    else {
        atomicOr(&W[w], v >> (s + n - 32));
        atomicOr(&W[w + 1], v << (64 - s - n));
    }
}

__device__ __forceinline__ uint32_t crc16_step(uint32_t c, uint32_t byte) {
    c ^= byte << 8;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) & 0xFFFFu : (c << 1) & 0xFFFFu;
    return c;
}
__device__ __forceinline__ uint32_t gf16_mulmod(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) {
        r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xFFFFu : (r << 1) & 0xFFFFu;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
__device__ __forceinline__ uint32_t crc8_bytes(const uint8_t* d, int n) {
    uint32_t c = 0;
    for (int i = 0; i < n; i++) {
        c ^= d[i];
        for (int k = 0; k < 8; k++) c = (c & 0x80u) ? ((c << 1) ^ 0x07u) & 0xFFu : (c << 1) & 0xFFu;
    }
    return c;
}

__device__ __forceinline__ int32_t load_sample(const uint8_t* pcm, uint64_t idx, int bytes) {
    const uint8_t* p = pcm + idx * bytes;
    if (bytes == 1) return (int32_t)p[0] - 128;                                   // WAV 8-bit is unsigned
    if (bytes == 2) return (int16_t)(p[0] | (p[1] << 8));
    return ((int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24)) >> 8;
}

__device__ __forceinline__ unsigned long long block_sum64(unsigned long long v, unsigned long long* scratch) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int i = 0; i < kT / 32; i++) t += scratch[i];
    return t;
}

}  // namespace

// dynamic smem: x[bs] int32 | u[bs] uint32
__global__ void __launch_bounds__(kFlacThreads) k_flac(const __grid_constant__ FlacArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    int32_t* x = reinterpret_cast<int32_t*>(smem_raw);
    uint32_t* u = reinterpret_cast<uint32_t*>(smem_raw) + A.block_size;
    __shared__ unsigned long long s_S[512];          // partition sums, level p at [2^p - 1 ...]
    __shared__ uint8_t s_k[512];                     // Rice parameter per partition, same indexing
    __shared__ unsigned long long s_cost[9];
    __shared__ unsigned int s_big[9];
    __shared__ unsigned long long s_scr[kT / 32];
    __shared__ unsigned long long s_scan[kT / 32];
    __shared__ unsigned long long s_fpos;
    __shared__ int s_order, s_bestp, s_mode;         // mode: 0 constant, 1 verbatim, 2 fixed
    __shared__ uint32_t s_crc[kT];

    const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t first = (uint64_t)blk * A.block_size;
    const int n = (int)min((uint64_t)A.block_size, A.n_samples - first);
    const int C = A.channels, bps = A.bits, bytes = A.bits >> 3;
    uint32_t* W = A.out + (size_t)blk * A.frame_words;
    const uint32_t vmask = bps == 32 ? 0xFFFFFFFFu : ((1u << bps) - 1u);

    // ---- frame header (read_frame_header_): thread 0
    if (tid == 0) {
        uint8_t h[16];
        int hn = 0;
        int bsc, bs_extra = 0;
        if (n == 192) bsc = 1;
        else if (n == 576 || n == 1152 || n == 2304 || n == 4608) bsc = 2 + (n == 1152) + 2 * (n == 2304) + 3 * (n == 4608);
        else if (n >= 256 && (n & (n - 1)) == 0 && n <= 32768) bsc = 8 + (31 - __clz(n)) - 8;
        else if (n <= 256) { bsc = 6; bs_extra = 8; }
        else { bsc = 7; bs_extra = 16; }
        int src = 0, sr_extra = 0;
        uint32_t sr_val = 0;
        const int r = A.sample_rate;
        const int std_r[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
        for (int i = 1; i < 12; i++) if (r == std_r[i]) src = i;
        if (!src) {
            if (r % 1000 == 0 && r / 1000 < 256) { src = 12; sr_extra = 8; sr_val = r / 1000; }
            else if (r < 65536) { src = 13; sr_extra = 16; sr_val = r; }
            else if (r % 10 == 0 && r / 10 < 65536) { src = 14; sr_extra = 16; sr_val = r / 10; }
        }
        const int ssc = bps == 8 ? 1 : bps == 12 ? 2 : bps == 16 ? 4 : bps == 20 ? 5 : bps == 24 ? 6 : 0;
        h[hn++] = 0xFF; h[hn++] = 0xF8;
        h[hn++] = (uint8_t)((bsc << 4) | src);
        h[hn++] = (uint8_t)(((C - 1) << 4) | (ssc << 1));
        const uint64_t fn = A.first_frame + (uint64_t)blk;                      // "UTF-8" coded frame number
        if (fn < 0x80) h[hn++] = (uint8_t)fn;
        else {
            const int nb = fn < 0x800 ? 2 : fn < 0x10000 ? 3 : fn < 0x200000 ? 4 : fn < 0x4000000 ? 5 : fn < 0x80000000ull ? 6 : 7;
            h[hn++] = (uint8_t)(((0xFF00u >> nb) & 0xFF) | (uint32_t)(fn >> (6 * (nb - 1))));
            for (int i = nb - 2; i >= 0; i--) h[hn++] = (uint8_t)(0x80 | ((fn >> (6 * i)) & 0x3F));
        }
        if (bs_extra == 8) h[hn++] = (uint8_t)(n - 1);
        if (bs_extra == 16) { h[hn++] = (uint8_t)((n - 1) >> 8); h[hn++] = (uint8_t)(n - 1); }
        if (sr_extra == 8) h[hn++] = (uint8_t)sr_val;
        if (sr_extra == 16) { h[hn++] = (uint8_t)(sr_val >> 8); h[hn++] = (uint8_t)sr_val; }
        h[hn] = (uint8_t)crc8_bytes(h, hn); hn++;
        for (int i = 0; i < hn; i++) put_bits(W, (uint64_t)i * 8, h[i], 8);
        s_fpos = (uint64_t)hn * 8;
    }
    __syncthreads();

    for (int c = 0; c < C; c++) {
        // ---- load the channel, signed
        for (int i = tid; i < n; i += kT) x[i] = load_sample(A.pcm, (first + i) * C + c, bytes);
        __syncthreads();
        const unsigned long long fpos = s_fpos;
        const int differs = __syncthreads_or(([&]() { int d = 0; for (int i = tid; i < n; i += kT) d |= x[i] != x[0]; return d; })());
        if (!differs) {                                                            // CONSTANT: 0|000000|0, value
            if (tid == 0) { put_bits(W, fpos, 0, 8); put_bits(W, fpos + 8, (uint32_t)x[0] & vmask, bps); s_fpos = fpos + 8 + bps; }
            __syncthreads();
            continue;
        }
        if (tid == 0) s_mode = 1;
        if (n >= 16) {
            // ---- fixed predictor search: sum |e_o| over i >= 4 for o = 0..4
            unsigned long long sm[5] = {0, 0, 0, 0, 0};
            for (int i = 4 + tid; i < n; i += kT) {
                const long long a0 = x[i], a1 = x[i - 1], a2 = x[i - 2], a3 = x[i - 3], a4 = x[i - 4];
                const long long e0 = a0, e1 = a0 - a1, e2 = a0 - 2 * a1 + a2, e3 = a0 - 3 * a1 + 3 * a2 - a3, e4 = a0 - 4 * a1 + 6 * a2 - 4 * a3 + a4;
                sm[0] += (unsigned long long)llabs(e0); sm[1] += (unsigned long long)llabs(e1); sm[2] += (unsigned long long)llabs(e2);
                sm[3] += (unsigned long long)llabs(e3); sm[4] += (unsigned long long)llabs(e4);
            }
            unsigned long long tot[5];
            for (int o = 0; o < 5; o++) tot[o] = block_sum64(sm[o], s_scr);
            int order = 0;
            for (int o = 1; o < 5; o++) if (tot[o] < tot[order]) order = o;
            // ---- zig-zag residual of the chosen order
            for (int i = order + tid; i < n; i += kT) {
                long long r;
                switch (order) {
                    case 0: r = x[i]; break;
                    case 1: r = (long long)x[i] - x[i - 1]; break;
                    case 2: r = (long long)x[i] - 2ll * x[i - 1] + x[i - 2]; break;
                    case 3: r = (long long)x[i] - 3ll * x[i - 1] + 3ll * x[i - 2] - x[i - 3]; break;
                    default: r = (long long)x[i] - 4ll * x[i - 1] + 6ll * x[i - 2] - 4ll * x[i - 3] + x[i - 4]; break;
                }
                const int32_t e = (int32_t)r;
                u[i] = ((uint32_t)e << 1) ^ (uint32_t)(e >> 31);
            }
            int pmax = 0;
            while (pmax < 8 && (n % (2 << pmax)) == 0 && (n >> (pmax + 1)) > order) pmax++;
            if (tid < 9) { s_cost[tid] = 0; s_big[tid] = 0; }
            __syncthreads();
            // ---- partition sums at the finest level (warp per partition), then the coarser levels by pairwise addition
            {
                const int np = 1 << pmax, m = n >> pmax;
                for (int j = warp; j < np; j += kT / 32) {
                    const int b = j ? j * m : order, end = (j + 1) * m;
                    unsigned long long s = 0;
                    for (int i = b + lane; i < end; i += 32) s += u[i];
#pragma unroll
                    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                    if (lane == 0) s_S[np - 1 + j] = s;
                }
                __syncthreads();
                for (int p = pmax - 1; p >= 0; p--) {
                    const int q = 1 << p;
                    for (int j = tid; j < q; j += kT) s_S[q - 1 + j] = s_S[2 * q - 1 + 2 * j] + s_S[2 * q - 1 + 2 * j + 1];
                    __syncthreads();
                }
            }
            // ---- Rice parameter per partition: largest k with (count << k) <= sum
            for (int idx = tid; idx < (2 << pmax) - 1; idx += kT) {
                const int p = 31 - __clz(idx + 1), j = idx + 1 - (1 << p), m = n >> p;
                const unsigned long long cnt = (unsigned long long)(j ? m : m - order), S = s_S[idx];
                int k = 0;
                while (k < 30 && (cnt << (k + 1)) <= S) k++;
                s_k[idx] = (uint8_t)k;
                if (k > 14) atomicOr(&s_big[p], 1u);
            }
            __syncthreads();
            // ---- exact bits of every partition order
            for (int p = 0; p <= pmax; p++) {
                const int np = 1 << p, m = n >> p;
                unsigned long long acc = 0;
                for (int j = warp; j < np; j += kT / 32) {
                    const int b = j ? j * m : order, end = (j + 1) * m, k = s_k[np - 1 + j];
                    unsigned long long s = 0;
                    for (int i = b + lane; i < end; i += 32) s += u[i] >> k;
#pragma unroll
                    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                    if (lane == 0) acc += s + (unsigned long long)(end - b) * (unsigned long long)(k + 1);
                }
                if (lane == 0 && acc) atomicAdd(&s_cost[p], acc);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long best = ~0ull;
                int bp = 0;
                for (int p = 0; p <= pmax; p++) {
                    const unsigned long long cst = s_cost[p] + (unsigned long long)(1 << p) * (s_big[p] ? 5 : 4);
                    if (cst < best) { best = cst; bp = p; }
                }
                s_order = order; s_bestp = bp;
                s_mode = (8ull + (unsigned long long)order * bps + 6 + best >= 8ull + (unsigned long long)n * bps) ? 1 : 2;
                s_cost[0] = best;
            }
        }
        __syncthreads();
        if (s_mode == 1) {                                                         // VERBATIM: 0|000001|0, n samples
            if (tid == 0) put_bits(W, fpos, 0x02, 8);
            for (int i = tid; i < n; i += kT) put_bits(W, fpos + 8 + (unsigned long long)i * bps, (uint32_t)x[i] & vmask, bps);
            __syncthreads();
            if (tid == 0) s_fpos = fpos + 8 + (unsigned long long)n * bps;
            __syncthreads();
            continue;
        }
        // ---- FIXED: 0|001ooo|0, warm-up, method, partition order, partitions
        const int order = s_order, bp = s_bestp, np = 1 << bp, m = n >> bp;
        const int rice2 = s_big[bp] ? 1 : 0, pbits = rice2 ? 5 : 4;
        unsigned long long pos0 = fpos;
        if (tid == 0) {
            put_bits(W, pos0, (uint32_t)((8 | order) << 1), 8);
            for (int i = 0; i < order; i++) put_bits(W, pos0 + 8 + (unsigned long long)i * bps, (uint32_t)x[i] & vmask, bps);
            put_bits(W, pos0 + 8 + (unsigned long long)order * bps, (uint32_t)rice2, 2);
            put_bits(W, pos0 + 10 + (unsigned long long)order * bps, (uint32_t)bp, 4);
        }
        pos0 += 14 + (unsigned long long)order * bps;
        // each thread owns a contiguous run of samples; block-wide exclusive scan of the code lengths
        const int per = (n - order + kT - 1) / kT;
        const int ib = order + tid * per, ie = min(ib + per, n);
        unsigned long long mine = 0;
        for (int i = ib; i < ie; i++) {
            const int j = i / m, k = s_k[np - 1 + j];
            mine += (u[i] >> k) + 1 + k;
            if (i == (j ? j * m : order)) mine += pbits;
        }
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        unsigned long long wbase = 0, total = 0;
        for (int i = 0; i < kT / 32; i++) { if (i < warp) wbase += s_scan[i]; total += s_scan[i]; }
        unsigned long long p = pos0 + wbase + incl - mine;
        for (int i = ib; i < ie; i++) {
            const int j = i / m, k = s_k[np - 1 + j];
            if (i == (j ? j * m : order)) { put_bits(W, p, (uint32_t)k, pbits); p += pbits; }
            const uint32_t q = u[i] >> k;
            p += q;                                                               // q zero bits: the buffer is zeroed
            put_bits(W, p, (1u << k) | (u[i] & ((1u << k) - 1u)), k + 1);         // stop bit + k low bits
            p += k + 1;
        }
        __syncthreads();
        if (tid == 0) s_fpos = pos0 + total;
        __syncthreads();
    }
    // ---- pad to a byte, big-endian words -> bytes, CRC-16 over the frame, append it
    __threadfence();
    __syncthreads();
    const unsigned long long fbits = s_fpos;
    const uint32_t nbytes = (uint32_t)((fbits + 7) >> 3);
    const uint32_t nwords = (nbytes + 3) >> 2;
    for (uint32_t i = tid; i < nwords; i += kT) { const uint32_t v = __ldcg(&W[i]); W[i] = __byte_perm(v, 0, 0x0123); }
    __threadfence();
    __syncthreads();
    const uint8_t* bytes_out = reinterpret_cast<const uint8_t*>(W);
    {
        const uint32_t L = (nbytes + kT - 1) / kT;
        const uint32_t b0 = min(nbytes, tid * L), b1 = min(nbytes, b0 + L);
        uint32_t crc = 0;
        for (uint32_t i = b0; i < b1; i++) crc = crc16_step(crc, __ldcg(reinterpret_cast<const unsigned char*>(bytes_out) + i));
        uint32_t after = nbytes - b1, pw = 0x0100u;                               // x^8
        while (after) {
            if (after & 1u) crc = gf16_mulmod(crc, pw);
            after >>= 1;
            if (after) pw = gf16_mulmod(pw, pw);
        }
        s_crc[tid] = crc;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0;
        for (int i = 0; i < kT; i++) c ^= s_crc[i];
        uint8_t* ob = reinterpret_cast<uint8_t*>(W);
        ob[nbytes] = (uint8_t)(c >> 8);
        ob[nbytes + 1] = (uint8_t)c;
        A.frame_len[blk] = nbytes + 2;
    }
}

cudaError_t launch_flac(const FlacArgs& a, int nblocks, cudaStream_t s) {
    const size_t smem = (size_t)a.block_size * 8;
    cudaError_t e = cudaFuncSetAttribute(k_flac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_flac<<<nblocks, kFlacThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace b200
