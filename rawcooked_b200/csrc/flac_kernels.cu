// sm_100a FLAC encoder kernel (kernels K5 + K6 fused): one CTA per block (= one FLAC frame).
//
// Bitstream = exact inverse of the reference's vendored libFLAC 1.3.2 decoder:
//   frame header      /root/reference/Source/Lib/ThirdParty/flac/src/libFLAC/stream_decoder.c:2159-2466 (read_frame_header_)
//   subframes         stream_decoder.c:2468-2540 (read_subframe_), fixed predictor fixed.c:336-395
//   partitioned Rice  stream_decoder.c:2745-2788 (read_residual_partitioned_rice_), bitreader.c:744
//   CRC-8 / CRC-16    crc.c:366-376; checked at stream_decoder.c:2075-2127
//   LPC subframes     stream_decoder.c:2627-2722 (read_subframe_lpc_), predictor arithmetic lpc.c:784-1100
// Per channel: constant test, fixed-predictor search (orders 0..4, criterion of fixed.c:217-273) against an LPC predictor of
// order 1..8 (integer window, exact integer autocorrelation, Levinson-Durbin in explicitly rounded double, 15-bit coefficients),
// zig-zag residual, exact search of the partition order 0..8 with Rice parameter floor(log2(mean)) per partition, VERBATIM
// fallback; then the bits:
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

// ---- LPC analysis: what one thread does between the block-wide passes. The decoder fixes the predictor
// arithmetic (lpc.c:784-1100, 64-bit sum whenever bps + precision + ilog2(order) > 32: stream_decoder.c:2710-2716) and the
// subframe syntax (read_subframe_lpc_, stream_decoder.c:2627-2722); the analysis is the encoder's own: integer Welch window,
// exact 64-bit integer autocorrelation (block-wide, order-independent), Levinson-Durbin in IEEE double where every operation
// is an explicitly rounded __d*_rn intrinsic (never contracted into a fused multiply-add, so a CPU computing the same
// operations one by one gets the same bits), quantisation with error feedback as in
// FLAC__lpc_quantize_coefficients (lpc.c:166-266).
constexpr int kLpcMaxOrder = 8;          // ffmpeg's level 5
constexpr int kLpcPrecision = 15;

struct LpcSet {
    int32_t q[kLpcMaxOrder][kLpcMaxOrder];
    int32_t shift[kLpcMaxOrder];
    int32_t valid[kLpcMaxOrder];
};

__device__ void lpc_levinson_quantise(const long long* ac, LpcSet* out) {
    for (int m = 0; m < kLpcMaxOrder; m++) out->valid[m] = 0;
    if (ac[0] <= 0) return;
    double R[kLpcMaxOrder + 1], a[kLpcMaxOrder], prev[kLpcMaxOrder];
    for (int l = 0; l <= kLpcMaxOrder; l++) R[l] = __ll2double_rn(ac[l]);
    double E = R[0];
    for (int m = 1; m <= kLpcMaxOrder; m++) {
        double acc = R[m];
        for (int j = 1; j < m; j++) { const double t = __dmul_rn(prev[j - 1], R[m - j]); acc = __dsub_rn(acc, t); }
        if (!(E > 0.0)) break;
        const double k = __ddiv_rn(acc, E);
        a[m - 1] = k;
        for (int j = 1; j < m; j++) { const double t = __dmul_rn(k, prev[m - 1 - j]); a[j - 1] = __dsub_rn(prev[j - 1], t); }
        { const double t = __dmul_rn(k, k); const double uu = __dsub_rn(1.0, t); E = __dmul_rn(E, uu); }
        for (int j = 0; j < m; j++) prev[j] = a[j];
        double cmax = 0.0;
        for (int j = 0; j < m; j++) { const double d = a[j] < 0.0 ? -a[j] : a[j]; if (d > cmax) cmax = d; }
        if (!(cmax > 0.0) || !(cmax < 1.0e9)) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(cmax);
        const int e = (int)((bits >> 52) & 0x7FF) - 1022;              // cmax = f * 2^e, f in [0.5, 1)
        int shift = kLpcPrecision - 1 - e;
        if (shift > 15) shift = 15;
        if (shift < 0) continue;
        const long long qmax = (1 << (kLpcPrecision - 1)) - 1, qmin = -(1 << (kLpcPrecision - 1));
        const double scale = (double)(1 << shift);
        double err = 0.0;
        out->shift[m - 1] = shift;
        for (int j = 0; j < m; j++) {
            const double t = __dmul_rn(a[j], scale);
            err = __dadd_rn(err, t);
            const double r = err >= 0.0 ? __dadd_rn(err, 0.5) : __dsub_rn(err, 0.5);
            long long q = __double2ll_rz(r);                            // round half away from zero
            if (q > qmax) q = qmax;
            if (q < qmin) q = qmin;
            err = __dsub_rn(err, __ll2double_rn(q));
            out->q[m - 1][j] = (int32_t)q;
        }
        out->valid[m - 1] = 1;
    }
}

// dynamic smem: x[bs] int32 | u[bs] uint32 | xw[bs] int32 (windowed samples of the LPC analysis)
__global__ void __launch_bounds__(kFlacThreads) k_flac(const __grid_constant__ FlacArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    int32_t* x = reinterpret_cast<int32_t*>(smem_raw);
    uint32_t* u = reinterpret_cast<uint32_t*>(smem_raw) + A.block_size;
    int32_t* xw = reinterpret_cast<int32_t*>(smem_raw) + 2 * A.block_size;
    __shared__ LpcSet s_lpc;
    __shared__ long long s_ac[kLpcMaxOrder + 1];
    __shared__ unsigned int s_lpc_bad;
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
        // ---- exact search of the partition order for the residual u[pred .. n): leaves the Rice parameters in s_k, the best
        // order in s_bestp and returns the bits of the residual section (without the 2 + 4 bits of its header); all threads
        auto rice_search = [&](int pred) -> unsigned long long {
            int pmax = 0;
            while (pmax < 8 && (n % (2 << pmax)) == 0 && (n >> (pmax + 1)) > pred) pmax++;
            if (tid < 9) { s_cost[tid] = 0; s_big[tid] = 0; }
            __syncthreads();
            {   // partition sums at the finest level (warp per partition), then the coarser levels by pairwise addition
                const int np = 1 << pmax, m = n >> pmax;
                for (int j = warp; j < np; j += kT / 32) {
                    const int b = j ? j * m : pred, end = (j + 1) * m;
                    unsigned long long sacc = 0;
                    for (int i = b + lane; i < end; i += 32) sacc += u[i];
#pragma unroll
                    for (int o = 16; o; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
                    if (lane == 0) s_S[np - 1 + j] = sacc;
                }
                __syncthreads();
                for (int p = pmax - 1; p >= 0; p--) {
                    const int q = 1 << p;
                    for (int j = tid; j < q; j += kT) s_S[q - 1 + j] = s_S[2 * q - 1 + 2 * j] + s_S[2 * q - 1 + 2 * j + 1];
                    __syncthreads();
                }
            }
            // Rice parameter per partition: largest k with (count << k) <= sum
            for (int idx = tid; idx < (2 << pmax) - 1; idx += kT) {
                const int p = 31 - __clz(idx + 1), j = idx + 1 - (1 << p), m = n >> p;
                const unsigned long long cntp = (unsigned long long)(j ? m : m - pred), S = s_S[idx];
                int k = 0;
                while (k < 30 && (cntp << (k + 1)) <= S) k++;
                s_k[idx] = (uint8_t)k;
                if (k > 14) atomicOr(&s_big[p], 1u);
            }
            __syncthreads();
            // exact bits of every partition order
            for (int p = 0; p <= pmax; p++) {
                const int np = 1 << p, m = n >> p;
                unsigned long long acc = 0;
                for (int j = warp; j < np; j += kT / 32) {
                    const int b = j ? j * m : pred, end = (j + 1) * m, k = s_k[np - 1 + j];
                    unsigned long long sacc = 0;
                    for (int i = b + lane; i < end; i += 32) sacc += u[i] >> k;
#pragma unroll
                    for (int o = 16; o; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
                    if (lane == 0) acc += sacc + (unsigned long long)(end - b) * (unsigned long long)(k + 1);
                }
                if (lane == 0 && acc) atomicAdd(&s_cost[p], acc);
            }
            __syncthreads();
            unsigned long long best = ~0ull;
            int bp = 0;
            for (int p = 0; p <= pmax; p++) {
                const unsigned long long cst = s_cost[p] + (unsigned long long)(1 << p) * (s_big[p] ? 5 : 4);
                if (cst < best) { best = cst; bp = p; }
            }
            __syncthreads();
            if (tid == 0) s_bestp = bp;
            __syncthreads();
            return best;
        };
        // residual of the quantised predictor of order m over x[m .. n), zig-zag folded into u; returns the block-wide sum of u
        // (every thread), s_lpc_bad != 0 if a residual does not fit 31 bits
        auto lpc_residual = [&](int m, bool store) -> unsigned long long {
            const int sh = s_lpc.shift[m - 1];
            long long qc[kLpcMaxOrder];
#pragma unroll
            for (int j = 0; j < kLpcMaxOrder; j++) qc[j] = j < m ? (long long)s_lpc.q[m - 1][j] : 0;
            unsigned long long part = 0;
            unsigned int bad = 0;
            for (int i = m + tid; i < n; i += kT) {
                long long sum = 0;
#pragma unroll
                for (int j = 0; j < kLpcMaxOrder; j++) if (j < m) sum += qc[j] * (long long)x[i - 1 - j];
                const int32_t pred = (int32_t)(sum >> sh);
                const long long e = (long long)x[i] - (long long)pred;
                if (e > 0x3FFFFFFFll || e < -0x40000000ll) bad = 1;
                const int32_t e32 = (int32_t)e;
                const uint32_t z = ((uint32_t)e32 << 1) ^ (uint32_t)(e32 >> 31);
                if (store) u[i] = z;
                part += z;
            }
            if (bad) atomicOr(&s_lpc_bad, 1u);
            return block_sum64(part, s_scr);
        };
        unsigned long long fixed_bits = ~0ull, lpc_bits = ~0ull;
        int order = 0, lpc_m = 0;
        if (n >= 16) {
            // ---- LPC candidate
            if (A.use_lpc && n > 2 * kLpcMaxOrder) {
                for (int i = tid; i < n; i += kT) {
                    const long long w15 = (4ll * i * (long long)(n - 1 - i) * 32767ll) / ((long long)(n - 1) * (long long)(n - 1));
                    xw[i] = (int32_t)(((long long)x[i] * w15) >> 15);
                }
                __syncthreads();
                long long acp[kLpcMaxOrder + 1];
#pragma unroll
                for (int l = 0; l <= kLpcMaxOrder; l++) acp[l] = 0;
                for (int i = tid; i < n; i += kT) {
                    const long long xi = xw[i];
#pragma unroll
                    for (int l = 0; l <= kLpcMaxOrder; l++) if (i >= l) acp[l] += xi * (long long)xw[i - l];
                }
#pragma unroll
                for (int l = 0; l <= kLpcMaxOrder; l++) {
                    const unsigned long long t = block_sum64((unsigned long long)acp[l], s_scr);     // two's complement: the sum wraps back
                    if (tid == 0) s_ac[l] = (long long)t;
                }
                __syncthreads();
                if (tid == 0) lpc_levinson_quantise(s_ac, &s_lpc);
                __syncthreads();
                unsigned long long best_est = ~0ull;
                for (int m = 1; m <= kLpcMaxOrder; m++) {
                    if (!s_lpc.valid[m - 1]) continue;                                                  // block-uniform
                    if (tid == 0) s_lpc_bad = 0;
                    __syncthreads();
                    const unsigned long long S = lpc_residual(m, false);
                    __syncthreads();
                    const uint32_t overflowed = s_lpc_bad;              // read by everyone before thread 0 clears it for the next order
                    __syncthreads();
                    if (overflowed) continue;
                    const unsigned long long cntm = (unsigned long long)(n - m);
                    int k = 0;
                    while (k < 30 && (cntm << (k + 1)) <= S) k++;
                    const unsigned long long est = cntm * (unsigned long long)(k + 1) + (S >> k) + (unsigned long long)m * (unsigned long long)(bps + kLpcPrecision);
                    if (est < best_est) { best_est = est; lpc_m = m; }
                }
                if (lpc_m) {
                    __syncthreads();
                    lpc_residual(lpc_m, true);
                    __syncthreads();
                    const unsigned long long cst = rice_search(lpc_m);
                    lpc_bits = 8ull + (unsigned long long)lpc_m * bps + 4 + 5 + (unsigned long long)lpc_m * kLpcPrecision + 6 + cst;
                }
            }
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
            for (int o = 1; o < 5; o++) if (tot[o] < tot[order]) order = o;
            __syncthreads();
            // ---- zig-zag residual of the chosen fixed order
            auto fixed_residual = [&]() {
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
                __syncthreads();
            };
            fixed_residual();
            fixed_bits = 8ull + (unsigned long long)order * bps + 6 + rice_search(order);
            const unsigned long long verbatim_bits = 8ull + (unsigned long long)n * bps;
            int mode = 1;
            if (lpc_bits < fixed_bits && lpc_bits < verbatim_bits) {
                mode = 3;
                __syncthreads();
                lpc_residual(lpc_m, true);                   // u and the Rice parameters of the winner again
                __syncthreads();
                rice_search(lpc_m);
            } else if (fixed_bits < verbatim_bits) {
                mode = 2;
            }
            if (tid == 0) { s_order = mode == 3 ? lpc_m : order; s_mode = mode; }
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
        // ---- FIXED: 0|001ooo|0, warm-up | LPC: 0|1ooooo|0, warm-up, precision, shift, coefficients; then method, partition
        // order, partitions
        const bool is_lpc = s_mode == 3;
        const int pred = s_order, bp = s_bestp, np = 1 << bp, m = n >> bp;
        const int rice2 = s_big[bp] ? 1 : 0, pbits = rice2 ? 5 : 4;
        unsigned long long pos0 = fpos;
        if (tid == 0) {
            put_bits(W, pos0, is_lpc ? (uint32_t)((32 | (pred - 1)) << 1) : (uint32_t)((8 | pred) << 1), 8);
            for (int i = 0; i < pred; i++) put_bits(W, pos0 + 8 + (unsigned long long)i * bps, (uint32_t)x[i] & vmask, bps);
            unsigned long long p = pos0 + 8 + (unsigned long long)pred * bps;
            if (is_lpc) {
                put_bits(W, p, (uint32_t)(kLpcPrecision - 1), 4); p += 4;
                put_bits(W, p, (uint32_t)s_lpc.shift[pred - 1], 5); p += 5;
                for (int j = 0; j < pred; j++) { put_bits(W, p, (uint32_t)s_lpc.q[pred - 1][j] & ((1u << kLpcPrecision) - 1u), kLpcPrecision); p += kLpcPrecision; }
            }
            put_bits(W, p, (uint32_t)rice2, 2);
            put_bits(W, p + 2, (uint32_t)bp, 4);
        }
        pos0 += 14 + (unsigned long long)pred * bps + (is_lpc ? 9ull + (unsigned long long)pred * kLpcPrecision : 0ull);
        // each thread owns a contiguous run of samples; block-wide exclusive scan of the code lengths
        const int per = (n - pred + kT - 1) / kT;
        const int ib = pred + tid * per, ie = min(ib + per, n);
        unsigned long long mine = 0;
        for (int i = ib; i < ie; i++) {
            const int j = i / m, k = s_k[np - 1 + j];
            mine += (u[i] >> k) + 1 + k;
            if (i == (j ? j * m : pred)) mine += pbits;
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
            if (i == (j ? j * m : pred)) { put_bits(W, p, (uint32_t)k, pbits); p += pbits; }
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
    const size_t smem = (size_t)a.block_size * 12;
    cudaError_t e = cudaFuncSetAttribute(k_flac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_flac<<<nblocks, kFlacThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace b200
