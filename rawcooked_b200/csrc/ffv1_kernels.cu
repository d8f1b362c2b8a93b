// sm_100a kernels of the B200 FFV1 encoder (see ffv1_kernels.cuh for the pipeline).
//
// Semantics = exact inverse of the reference decoder:
//   sample order / borders  /root/reference/Source/Lib/CoDec/FFV1/FFV1_Slice.cpp:406-444 (LineThenPlane), :447-472 (Line)
//   predictor / context     FFV1_Slice.cpp:21-93
//   binarisation            Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:135-171 (rangecoder::s)
//   bin arithmetic          FFV1_RangeCoder.cpp:71-102 (rangecoder::b), byte accounting :51-56
//   state reset per frame   Source/Lib/CoDec/FFV1/Coder/FFV1_Coder_RangeCoder.cpp:25-48
//   footer                  FFV1_Slice.cpp:247-253, :301-315; FFV1_Frame.cpp:177-197
//   pixel layouts + RCT     Source/Lib/Transform/Transform.cpp:29-37, :70-420
#include "ffv1_kernels.cuh"

#include "../../include/b200enc.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------------------------
// pixel fetch: the three stored components of pixel x in a payload row (value bits only)
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

__device__ __forceinline__ void load_rgb(const uint8_t* __restrict__ row, int layout, int x, int& r, int& g, int& b) {
    switch (layout) {
        case B200_DPX_RGB_8: case B200_TIFF_RGB_8: {
            const uint8_t* q = row + 3 * x;
            r = q[0]; g = q[1]; b = q[2];
            break;
        }
        case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: {
            uint32_t v = *reinterpret_cast<const uint32_t*>(row + 4 * x);
            if (layout == B200_DPX_RGB_10_FILLED_A_BE) v = bswap32(v);
            r = (v >> 22) & 1023; g = (v >> 12) & 1023; b = (v >> 2) & 1023;
            break;
        }
        case B200_DPX_RGB_12_FILLED_A_LE: case B200_DPX_RGB_16_LE: case B200_TIFF_RGB_16_LE: {
            const uint16_t* q = reinterpret_cast<const uint16_t*>(row + 6 * x);
            int sh = layout == B200_DPX_RGB_12_FILLED_A_LE ? 4 : 0;
            r = q[0] >> sh; g = q[1] >> sh; b = q[2] >> sh;
            break;
        }
        case B200_DPX_RGB_12_FILLED_A_BE: case B200_DPX_RGB_16_BE: case B200_TIFF_RGB_16_BE: {
            const uint16_t* q = reinterpret_cast<const uint16_t*>(row + 6 * x);
            int sh = layout == B200_DPX_RGB_12_FILLED_A_BE ? 4 : 0;
            uint32_t a = q[0], c = q[1], d = q[2];
            r = (int)(__byte_perm(a, 0, 0x4401)) >> sh;
            g = (int)(__byte_perm(c, 0, 0x4401)) >> sh;
            b = (int)(__byte_perm(d, 0, 0x4401)) >> sh;
            break;
        }
        case B200_DPX_RGB_12_PACKED_BE: {
            // component k = 3x + c lives at bit 12k (LSB first) of the row seen as big-endian 32-bit words
            const uint32_t* wds = reinterpret_cast<const uint32_t*>(row);
            int v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                uint32_t bit = (uint32_t)(3 * x + c) * 12u;
                uint32_t wi = bit >> 5, sh = bit & 31;
                uint32_t lo = bswap32(wds[wi]);
                uint32_t hi = sh > 20 ? bswap32(wds[wi + 1]) : 0u;   // 12 bits straddle only when sh > 20
                v[c] = (int)(__funnelshift_r(lo, hi, sh) & 0xFFFu);
            }
            r = v[0]; g = v[1]; b = v[2];
            break;
        }
        default: r = g = b = 0;
    }
}

__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

// ------------------------------------------------------------------------------------------------------------------
// k_model
//
// shared memory carve-up (dynamic):
//   states  nctx*32 B      adaptive state of every (context, slot) of this plane-set
//   ring    3 rows x planes x wmax int32   RCT'd samples of rows y, y-1, y-2
//   s_val   wmax int32     folded residual of the plane-row being coded (after the context-sign flip)
//   s_off   wmax uint32    first bin of each sample inside the row
//   s_ctx   wmax uint16    context index (>= 0)
//   qtab 5*256 int16, trans 512 B, warp totals
struct ModelSmem {
    uint8_t* states; int32_t* ring; int32_t* val; uint32_t* off; uint16_t* ctx; int16_t* qtab; uint8_t* trans; uint32_t* wtot;
};
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

size_t model_smem_bytes(int nctx, int sstride, int wmax, int planes) {
    size_t n = align16((size_t)nctx * sstride);
    n += align16((size_t)3 * planes * wmax * 4);
    n += align16((size_t)wmax * 4) * 2;
    n += align16((size_t)wmax * 2);
    n += 5 * 256 * 2 + 512 + 32 * 4;
    return n;
}

__device__ __forceinline__ ModelSmem carve(uint8_t* base, int nctx, int sstride, int wmax, int planes) {
    ModelSmem m;
    m.states = base; base += align16((size_t)nctx * sstride);
    m.ring = reinterpret_cast<int32_t*>(base); base += align16((size_t)3 * planes * wmax * 4);
    m.val = reinterpret_cast<int32_t*>(base); base += align16((size_t)wmax * 4);
    m.off = reinterpret_cast<uint32_t*>(base); base += align16((size_t)wmax * 4);
    m.ctx = reinterpret_cast<uint16_t*>(base); base += align16((size_t)wmax * 2);
    m.qtab = reinterpret_cast<int16_t*>(base); base += 5 * 256 * 2;
    m.trans = base; base += 512;
    m.wtot = reinterpret_cast<uint32_t*>(base);
    return m;
}

__global__ void __launch_bounds__(kModelThreads, 1) k_model(const __grid_constant__ EncArgs A, int band) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int slice = blockIdx.x >> 1, ps = blockIdx.x & 1, frame = blockIdx.y;
    const SliceGeom g = A.geom[slice];
    const int r0 = band * A.band_rows;
    if (r0 >= g.h) return;
    const int r1 = min(r0 + A.band_rows, g.h);
    const int planes = ps ? 2 : 1;
    const int w = g.w, wmax = A.wmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kModelThreads / 32;
    ModelSmem S = carve(smem_raw, A.nctx, A.sstride, wmax, planes);

    const size_t fs = (size_t)frame * A.nslices + slice;
    const int state_bytes = (int)align16((size_t)A.nctx * A.sstride);
    uint8_t* save = A.state_save + (fs * 2 + ps) * (size_t)state_bytes;
    // 8-bit streams never use slots 10, 20, 21, 30, 31 (e <= 8): 27 states per context so the large model fits in smem
    const int slot = A.sstride == 32 ? lane : lane - (lane > 10) - 2 * (lane > 21);
    {   // context states: 128 at the start of every frame (intra-only), else carried from the previous band
        const int n16 = state_bytes >> 4;
        uint4* d = reinterpret_cast<uint4*>(S.states);
        if (band == 0) {
            const uint4 v = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
            for (int i = tid; i < n16; i += kModelThreads) d[i] = v;
        } else {
            const uint4* s = reinterpret_cast<const uint4*>(save);
            for (int i = tid; i < n16; i += kModelThreads) d[i] = s[i];
        }
        for (int i = tid; i < 5 * 256; i += kModelThreads) S.qtab[i] = A.qtab[i];
        for (int i = tid; i < 512; i += kModelThreads) S.trans[i] = A.trans[i];
    }
    const uint8_t* fin = A.in + (size_t)frame * A.frame_bytes;
    const int off = 1 << A.bits;

    // RCT one payload row into ring slot (y+3)%3; rows above the slice are zero (FFV1_Slice.cpp:409-410 memset)
    auto load_row = [&](int y) {
        int32_t* dst = S.ring + (size_t)((y + 3) % 3) * planes * wmax;
        if (y < 0) {
            for (int i = tid; i < planes * wmax; i += kModelThreads) dst[i] = 0;
            return;
        }
        const uint8_t* row = fin + (size_t)(g.y0 + y) * A.row_bytes;
        for (int x = tid; x < w; x += kModelThreads) {
            int r, gg, b;
            load_rgb(row, A.layout, g.x0 + x, r, gg, b);
            if (A.swap_bg) { int t = gg; gg = b; b = t; }
            b -= gg; r -= gg; gg += (b + r) >> 2; b += off; r += off;
            if (ps == 0) dst[x] = gg;
            else { dst[x] = b; dst[wmax + x] = r; }
        }
    };
    load_row(r0 - 2);
    load_row(r0 - 1);

    uint16_t* bins = (ps ? A.binsC : A.binsY) + fs * (ps ? A.capC : A.capY);
    uint32_t pos = 0;                       // bins emitted so far in this band (CTA-uniform)
    const int K = (w + kModelThreads - 1) / kModelThreads;
    const int sbits = A.sbits;

    for (int y = r0; y < r1; y++) {
        load_row(y);
        __syncthreads();
        uint32_t row_bins = 0;
        for (int pl = 0; pl < planes; pl++) {
            const int32_t* cur = S.ring + ((size_t)((y + 3) % 3) * planes + pl) * wmax;
            const int32_t* prv = S.ring + ((size_t)((y + 2) % 3) * planes + pl) * wmax;
            const int32_t* pp2 = S.ring + ((size_t)((y + 1) % 3) * planes + pl) * wmax;
            // ---- K2: prediction, context, fold; per-thread bin counts
            uint32_t mine = 0;
            const int xb = tid * K, xe = min(xb + K, w);
            for (int x = xb; x < xe; x++) {
                const int T = prv[x];
                const int RT = prv[x + 1 < w ? x + 1 : w - 1];            // sample[1][w] = sample[1][w-1]
                const int L = x > 0 ? cur[x - 1] : prv[0];                 // sample[0][-1] = sample[1][0]
                const int LT = x > 0 ? prv[x - 1] : pp2[0];                // what sample[1][-1] was set to one row earlier
                int ctx = S.qtab[(L - LT) & 255] + S.qtab[256 + ((LT - T) & 255)] + S.qtab[512 + ((T - RT) & 255)];
                if (A.is5) {
                    const int LL = x > 1 ? cur[x - 2] : (x == 1 ? prv[0] : 0);
                    const int TT = pp2[x];
                    ctx += S.qtab[768 + ((LL - L) & 255)] + S.qtab[1024 + ((TT - T) & 255)];
                }
                int d = cur[x] - median3(L, L + T - LT, T);
                if (ctx < 0) { ctx = -ctx; d = -d; }
                d = (d << (32 - sbits)) >> (32 - sbits);                  // fold to sbits, sign-extended
                S.val[x] = d;
                S.ctx[x] = (uint16_t)ctx;
                mine += d ? 2 * (31 - __clz(abs(d))) + 3 : 1;
            }
            // ---- exclusive scan of bin counts over the row
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) S.wtot[warp] = incl;
            __syncthreads();
            uint32_t wv = lane < NW ? S.wtot[lane] : 0, wincl = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o); if (lane >= o) wincl += t; }
            const uint32_t total = __shfl_sync(0xffffffffu, wincl, NW - 1);
            const uint32_t wbase = __shfl_sync(0xffffffffu, wincl - wv, warp);
            uint32_t o = wbase + incl - mine;
            for (int x = xb; x < xe; x++) {
                S.off[x] = o;
                const int d = S.val[x];
                o += d ? 2 * (31 - __clz(abs(d))) + 3 : 1;
            }
            __syncthreads();
            // ---- K3: adaptive state seen by every bin. Warp `warp` owns the contexts with (ctx mod NW) == warp and walks
            // its samples in bitstream order; lane s owns slot s of the 32-state context (slots are independent chains).
            uint16_t* out = bins + pos;
            for (int c0 = 0; c0 < w; c0 += 32) {
                const int xm = c0 + lane;
                const uint32_t myctx = xm < w ? S.ctx[xm] : 0xFFFFu;
                uint32_t todo = __ballot_sync(0xffffffffu, xm < w && (int)(myctx % NW) == warp);
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t cx = __shfl_sync(0xffffffffu, myctx, j);
                    const int v = S.val[c0 + j];
                    const uint32_t ob = S.off[c0 + j];
                    // which bins of this symbol live in slot `lane` (rangecoder::s, FFV1_RangeCoder.cpp:135-171)
                    const uint32_t a = (uint32_t)abs(v);
                    const int e = 31 - __clz(a | 1);
                    int n = 0, i0 = 0, step = 0;     // n bins; bin k uses index i = i0 + k*step
                    if (lane == 0) n = 1;
                    else if (v != 0) {
                        if (lane <= 9) { n = (lane - 1) <= e; i0 = lane - 1; }
                        else if (lane == 10) { n = e >= 9 ? e - 8 : 0; i0 = 9; step = 1; }
                        else if (lane <= 21) { n = (lane - 11) == min(e, 10); }
                        else if (lane <= 30) { n = (lane - 22) < e; i0 = lane - 22; }
                        else { n = e > 9 ? e - 9 : 0; i0 = e - 1; step = -1; }
                    }
                    if (n) {
                        uint8_t* sp = S.states + cx * A.sstride + slot;
                        uint32_t st = *sp;
                        for (int k = 0; k < n; k++) {
                            const int i = i0 + k * step;
                            uint32_t bit, idx;
                            if (lane == 0) { bit = v == 0; idx = 0; }
                            else if (lane <= 10) { bit = i < e; idx = 1 + i; }
                            else if (lane <= 21) { bit = v < 0; idx = 2 * e + 2; }
                            else { bit = (a >> i) & 1; idx = 2 * e + 1 - i; }
                            out[ob + idx] = (uint16_t)(st | (bit << 8));
                            st = S.trans[(bit << 8) | st];
                        }
                        *sp = (uint8_t)st;
                    }
                }
            }
            pos += total;
            row_bins += total;
            __syncthreads();
        }
        if (tid == 0) A.rowcnt[(fs * A.band_rows + (y - r0)) * 2 + ps] = row_bins;
    }
    if (r1 < g.h) {   // carry the states to the next band
        const int n16 = state_bytes >> 4;
        const uint4* s = reinterpret_cast<const uint4*>(S.states);
        uint4* d = reinterpret_cast<uint4*>(save);
        for (int i = tid; i < n16; i += kModelThreads) d[i] = s[i];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_code: one lane per (frame, slice)
struct Coder {
    uint32_t low, range;
    int32_t pending;
    uint32_t run, pos, crc;
    uint8_t* out;
    uint32_t cap;
    const uint32_t* crct;
    bool overflow;

    __device__ __forceinline__ void put(uint32_t b) {
        if (pos < cap) out[pos] = (uint8_t)b; else overflow = true;
        pos++;
        crc = (crc << 8) ^ crct[(crc >> 24) ^ (b & 255u)];
    }
    __device__ __forceinline__ void shift() {      // one renormalisation step (range < 0x100 on entry)
        if (pending < 0) pending = (int32_t)(low >> 8);
        else if (low <= 0xFF00u) { put((uint32_t)pending); for (; run; run--) put(0xFFu); pending = (int32_t)(low >> 8); }
        else if (low >= 0x10000u) { put((uint32_t)pending + 1); for (; run; run--) put(0u); pending = (int32_t)((low >> 8) & 255u); }
        else run++;
        low = (low & 255u) << 8;
        range <<= 8;
    }
    __device__ __forceinline__ void bin(uint32_t rec) {
        const uint32_t st = rec & 255u;
        const uint32_t r1 = (range * st) >> 8;
        if (rec & 256u) { low += range - r1; range = r1; }
        else range -= r1;
        if (range < 0x100u) shift();
    }
};

__global__ void __launch_bounds__(32) k_code(const __grid_constant__ EncArgs A, int band, int nframes) {
    __shared__ uint32_t s_crc[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = A.crc_table[i];
    __syncthreads();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nframes * A.nslices) return;
    const int slice = gid % A.nslices;
    const SliceGeom g = A.geom[slice];
    const int r0 = band * A.band_rows;
    if (r0 >= g.h) return;
    const int r1 = min(r0 + A.band_rows, g.h);

    Coder c;
    c.out = A.scratch + (size_t)gid * A.slice_cap;
    c.cap = (uint32_t)A.slice_cap - 8;
    c.crct = s_crc;
    c.overflow = false;
    if (band == 0) {
        c.low = 0; c.range = 0xFF00u; c.pending = -1; c.run = 0; c.pos = 0; c.crc = 0;
        const uint16_t* hb = A.hdr_bins + (size_t)slice * kMaxHeaderBins;
        const int nh = A.hdr_cnt[slice];
        for (int i = 0; i < nh; i++) c.bin(hb[i]);
    } else {
        const CoderState s = A.cstate[gid];
        c.low = s.low; c.range = s.range; c.pending = s.pending; c.run = s.run; c.pos = s.pos; c.crc = s.crc;
    }
    const uint16_t* by = A.binsY + (size_t)gid * A.capY;
    const uint16_t* bc = A.binsC + (size_t)gid * A.capC;
    const uint32_t* rc = A.rowcnt + (size_t)gid * A.band_rows * 2;
    uint64_t nb = 0;
    for (int y = r0; y < r1; y++) {
        const uint32_t ny = rc[(y - r0) * 2], nc = rc[(y - r0) * 2 + 1];
        for (uint32_t i = 0; i < ny; i++) c.bin(by[i]);
        by += ny;
        for (uint32_t i = 0; i < nc; i++) c.bin(bc[i]);
        bc += nc;
        nb += ny + nc;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(A.flags + 2), (unsigned long long)nb);
    if (r1 == g.h) {
        c.bin(129u);                                           // terminator bin, state 129, bit 0 (FFV1_Slice.cpp:334-343)
        c.range = 0xFFu; c.low += 0xFFu; c.shift();            // flush so that BytesUsed() == payload size (:297-299)
        c.range = 0xFFu; c.shift();
        const uint32_t n = c.pos;
        c.cap += 8;
        c.put(n >> 16); c.put(n >> 8); c.put(n);               // slice_size
        if (A.ec) {
            c.put(0);                                          // error_status
            const uint32_t crc = c.crc;                        // parity: CRC of the whole slice becomes 0
            c.put(crc >> 24); c.put(crc >> 16); c.put(crc >> 8); c.put(crc);
        }
        A.slice_size[gid] = c.pos;
    } else {
        CoderState s;
        s.low = c.low; s.range = c.range; s.pending = c.pending; s.run = c.run; s.pos = c.pos; s.crc = c.crc; s.offY = 0; s.offC = 0;
        A.cstate[gid] = s;
    }
    if (c.overflow) atomicOr(A.flags, 1u);
}

// ------------------------------------------------------------------------------------------------------------------
// k_scan: slice sizes -> arena offsets (single CTA; the arrays are tiny)
__global__ void __launch_bounds__(1024) k_scan(const __grid_constant__ EncArgs A, int nframes) {
    __shared__ uint64_t s_part[1024];
    const int n = nframes * A.nslices;
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(b + per, n);
    uint64_t sum = 0;
    for (int i = b; i < e; i++) sum += A.slice_size[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint64_t t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    uint64_t off = s_part[threadIdx.x] - sum;
    for (int i = b; i < e; i++) {
        A.slice_off[i] = off;
        off += A.slice_size[i];
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nframes; f += 1024) {
        const int first = f * A.nslices, last = first + A.nslices - 1;
        A.frame_off[f] = A.slice_off[first];
        A.frame_len[f] = A.slice_off[last] + A.slice_size[last] - A.slice_off[first];
    }
    if (threadIdx.x == 1023 && s_part[1023] > A.arena_cap) atomicOr(A.flags, 2u);
}

// k_pack: copy each slice from its scratch region to its place in the arena (16-byte stores, funnel-shifted loads)
__global__ void __launch_bounds__(256) k_pack(const __grid_constant__ EncArgs A) {
    const int gid = blockIdx.x;
    const uint32_t n = A.slice_size[gid];
    const uint64_t off = A.slice_off[gid];
    if (off + n > A.arena_cap) return;
    const uint8_t* src = A.scratch + (size_t)gid * A.slice_cap;
    uint8_t* dst = A.arena + off;
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    const int t = blockIdx.y * blockDim.x + threadIdx.x, nt = gridDim.y * blockDim.x;
    if (t < (int)head) dst[t] = src[t];
    const uint32_t body = (n - head) >> 4;
    const uint32_t sh = (head & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = t; i < body; i += nt) {
        const uint32_t* p = sw + i * 4;
        uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = sh ? p[4] : 0;
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
        dv[i] = v;
    }
    const uint32_t tail0 = head + (body << 4);
    if (t < (int)(n - tail0)) dst[tail0 + t] = src[tail0 + t];
}

// ------------------------------------------------------------------------------------------------------------------
cudaError_t configure_kernels(int nctx, int sstride, int wmax) {
    size_t need = model_smem_bytes(nctx, sstride, wmax, 2);
    return cudaFuncSetAttribute(k_model, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
}

cudaError_t launch_model(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    dim3 grid(a.nslices * 2, nframes);
    size_t smem = model_smem_bytes(a.nctx, a.sstride, a.wmax, 2);
    k_model<<<grid, kModelThreads, smem, s>>>(a, band);
    return cudaGetLastError();
}

cudaError_t launch_code(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    int n = nframes * a.nslices;
    k_code<<<(n + 31) / 32, 32, 0, s>>>(a, band, nframes);
    return cudaGetLastError();
}

cudaError_t launch_pack(const EncArgs& a, int nframes, cudaStream_t s) {
    k_scan<<<1, 1024, 0, s>>>(a, nframes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid(nframes * a.nslices, 4);
    k_pack<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace b200
