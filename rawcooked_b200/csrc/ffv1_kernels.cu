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

constexpr uint32_t kNopRec = 0x100u;

__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

// ------------------------------------------------------------------------------------------------------------------
// k_model
//
// One CTA per (frame, slice, plane-set) and band. Shared memory (dynamic):
//   states  nctx*sstride B   adaptive state of every (context, slot) of this plane-set
//   ring    3 rows x planes x wmax int32   RCT'd samples of rows y, y-1, y-2
//   val/off/ctx per sample of the plane-row being coded: folded residual (after the context-sign flip), first record of the
//           sample inside the row, context index (>= 0)
//   cnt8    per-context occurrence counter inside the current plane-row (touched only by the warp owning the context)
//   tmp8    per-sample scratch byte: position inside its chunk-class group, later the round of the sample
//   plist   samples of the row partitioned by owner warp (ctx mod 16), each part in x order
//   ccnt    [chunk][class] counts -> exclusive prefix over chunks; ctot/coff the same for record counts
//   rlist   kRounds lists of samples: round r = samples whose context occurred r times earlier in the row
//   stage   the row's records, copied out coalesced at the end of the row
constexpr int kRounds = 2;
constexpr int kMaxChunks = 64;            // wmax <= 2048
constexpr int kSlutRow = 36;              // slot of bin b of a symbol with exponent e: slut[e * kSlutRow + b], e <= 16, b <= 34
constexpr int kSlutBytes = 17 * kSlutRow + 12;   // 624, keeps 16-byte alignment
struct ModelSmem {
    uint8_t* states; int32_t* ring; int32_t* val; uint32_t* off; uint16_t* ctx; uint8_t* cnt8; uint8_t* tmp8; uint16_t* plist;
    uint16_t* ccnt; uint32_t* ctot; uint16_t* clstot; uint16_t* cstart; uint16_t* rlist; uint32_t* rfill; uint32_t* misc;
    uint16_t* stage; int16_t* qtab; uint8_t* trans; uint8_t* slut; uint32_t* t2;
};
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline int rlist_entries(int wmax) {  // NOLINT
    int n = 0;
    for (int r = 0; r < kRounds; r++) n += wmax / (r + 1) + 1;
    return n;
}
// bytes of everything except the record staging area
size_t model_smem_fixed(int nctx, int sstride, int wmax, int planes) {
    size_t n = align16((size_t)nctx * sstride);
    n += align16((size_t)3 * planes * wmax * 4);
    n += align16((size_t)wmax * 4) * 2;          // val, off
    n += align16((size_t)wmax * 2) * 2;          // ctx, plist
    n += align16((size_t)nctx);                  // cnt8
    n += align16((size_t)wmax);                  // tmp8
    n += kMaxChunks * kModelWarps * 2 + kMaxChunks * 4 + 32 * 2 + 32 * 2;   // ccnt, ctot, clstot, cstart
    n += align16((size_t)rlist_entries(wmax) * 2);
    n += 16 * 4 + 16 * 4;                        // rfill, misc
    n += 5 * 256 * 2 + 512 + kSlutBytes + 512 * 4;
    return n;
}
size_t model_smem_bytes(int nctx, int sstride, int wmax, int planes, int stage_cap) {
    return model_smem_fixed(nctx, sstride, wmax, planes) + align16((size_t)stage_cap * 2);
}

__device__ __forceinline__ ModelSmem carve(uint8_t* base, int nctx, int sstride, int wmax, int planes) {
    ModelSmem m;
    m.states = base; base += align16((size_t)nctx * sstride);
    m.ring = reinterpret_cast<int32_t*>(base); base += align16((size_t)3 * planes * wmax * 4);
    m.val = reinterpret_cast<int32_t*>(base); base += align16((size_t)wmax * 4);
    m.off = reinterpret_cast<uint32_t*>(base); base += align16((size_t)wmax * 4);
    m.ctx = reinterpret_cast<uint16_t*>(base); base += align16((size_t)wmax * 2);
    m.plist = reinterpret_cast<uint16_t*>(base); base += align16((size_t)wmax * 2);
    m.cnt8 = base; base += align16((size_t)nctx);
    m.tmp8 = base; base += align16((size_t)wmax);
    m.ccnt = reinterpret_cast<uint16_t*>(base); base += kMaxChunks * kModelWarps * 2;
    m.ctot = reinterpret_cast<uint32_t*>(base); base += kMaxChunks * 4;
    m.clstot = reinterpret_cast<uint16_t*>(base); base += 32 * 2;
    m.cstart = reinterpret_cast<uint16_t*>(base); base += 32 * 2;
    m.rlist = reinterpret_cast<uint16_t*>(base); base += align16((size_t)rlist_entries(wmax) * 2);
    m.rfill = reinterpret_cast<uint32_t*>(base); base += 16 * 4;
    m.misc = reinterpret_cast<uint32_t*>(base); base += 16 * 4;
    m.qtab = reinterpret_cast<int16_t*>(base); base += 5 * 256 * 2;
    m.trans = base; base += 512;
    m.slut = base; base += kSlutBytes;
    m.t2 = reinterpret_cast<uint32_t*>(base); base += 512 * 4;
    m.stage = reinterpret_cast<uint16_t*>(base);
    return m;
}

// bin record handed from k_model to k_range / k_emit: sp | bit << 9 where sp = bit ? state : 256 - state is the 8-bit
// probability of the coded value. kNopRec (sp = 256, bit 0) leaves the coder untouched; it pads every plane-row to whole
// 128-byte blocks.
__device__ __forceinline__ uint32_t make_rec(uint32_t st, uint32_t bit) { return (bit ? st : 256u - st) | (bit << 9); }

constexpr int kMaxPixPerThread = 2048 / kModelThreads;      // wmax <= 2048

// -DB200_PHASE_TIMING: thread 0 of every CTA accumulates the cycles between phase boundaries into flags[16 + 2*phase]
#ifdef B200_PHASE_TIMING
#define PHASE_MARK(k) do { if (tid == 0) { long long t_ = clock64(); ph[k] += t_ - tlast; tlast = t_; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

// Byte offset of (context, slot) in the shared-memory state table. With 32 states per context every row starts on the
// same 8 banks as the row four contexts further, so the word inside the row is XOR-swizzled with context bits 2..4:
// lanes that touch the same slot of different contexts then spread over all 32 banks.
__device__ __forceinline__ uint32_t state_off(uint32_t ctx, uint32_t slot, bool compact) {
    return compact ? ctx * 27u + slot : ctx * 32u + (slot ^ (((ctx >> 2) & 7u) << 2));
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
    const uint32_t lt = (1u << lane) - 1u;
    constexpr int NW = kModelThreads / 32;
    static_assert(NW == 16 || NW == 32, "context class = ctx & (NW - 1) = owner warp");
    ModelSmem S = carve(smem_raw, A.nctx, A.sstride, wmax, planes);
    const uint32_t stage_cap = (uint32_t)A.stage_cap;
    const bool compact = A.sstride != 32;   // 8-bit streams never use slots 10, 20, 21, 30, 31 (e <= 8): 27 states per context

    const size_t fs = (size_t)frame * A.nslices + slice;
    const int state_bytes = (int)align16((size_t)A.nctx * A.sstride);
    uint8_t* save = A.state_save + (fs * 2 + ps) * (size_t)state_bytes;
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
        for (int i = tid; i < (int)align16((size_t)A.nctx); i += kModelThreads) S.cnt8[i] = 0;
        // t2[bit << 8 | state] = next state | record << 16: one lookup per bin gives the transition and the coder record
        for (int i = tid; i < 512; i += kModelThreads) {
            const uint32_t st = i & 255, bit = i >> 8;
            S.t2[i] = (uint32_t)A.trans[i] | (make_rec(st, bit) << 16);
        }
        // slot of bin b of a symbol with exponent e (rangecoder::s, FFV1_RangeCoder.cpp:135-171): zero flag (slot 0), unary
        // exponent (1..10), mantissa from the top bit down (22..31), sign (11..21)
        for (int i = tid; i < 17 * kSlutRow; i += kModelThreads) {
            const int e = i / kSlutRow, b = i % kSlutRow;
            int slot = b <= e + 1 ? min(b, 10) : (b <= 2 * e + 1 ? 22 + min(2 * e + 1 - b, 9) : 11 + min(e, 10));
            if (compact) slot = slot - (slot > 10) - 2 * (slot > 21);
            S.slut[i] = (uint8_t)slot;
        }
    }
    const uint8_t* fin = A.in + (size_t)frame * A.frame_bytes;
    const int off = 1 << A.bits;

    // forward RCT of pixel x of slice row y (inverse of Transform.cpp:29-37); p0 = Y or Cb, p1 = Cr
    auto fetch = [&](int y, int x, int& p0, int& p1) {
        const uint8_t* row = fin + (size_t)(g.y0 + y) * A.row_bytes;
        int r, gg, b;
        load_rgb(row, A.layout, g.x0 + x, r, gg, b);
        if (A.swap_bg) { int t = gg; gg = b; b = t; }
        b -= gg; r -= gg; gg += (b + r) >> 2; b += off; r += off;
        if (ps == 0) { p0 = gg; p1 = 0; } else { p0 = b; p1 = r; }
    };
    // rows above the slice are zero (FFV1_Slice.cpp:409-410 memset)
    auto load_row = [&](int y) {
        int32_t* dst = S.ring + (size_t)((y + 3) % 3) * planes * wmax;
        if (y < 0) {
            for (int i = tid; i < planes * wmax; i += kModelThreads) dst[i] = 0;
            return;
        }
        for (int x = tid; x < w; x += kModelThreads) {
            int p0, p1;
            fetch(y, x, p0, p1);
            dst[x] = p0;
            if (ps) dst[wmax + x] = p1;
        }
    };
    load_row(r0 - 2);
    load_row(r0 - 1);
    load_row(r0);

    uint16_t* bins = (ps ? A.binsC : A.binsY) + fs * (ps ? A.capC : A.capY);
    uint32_t pos = 0;                       // records emitted so far in this band (CTA-uniform; segments start on multiples of 64)
    uint32_t seg_extra = 0;                 // records already in the current segment (slice header ahead of the first Y row)
    if (band == 0 && ps == 0) {
        const int nh = A.hdr_cnt[slice];
        for (int i = tid; i < nh; i += kModelThreads) bins[i] = A.hdr_bins[(size_t)slice * kMaxHeaderBins + i];
        pos = seg_extra = (uint32_t)nh;
    }
    __syncthreads();

    const int nchunk = (w + 31) >> 5;
    const int KC = (nchunk + NW - 1) / NW;  // 32-sample chunks per warp
    const int sbits = A.sbits;
    uint32_t* rl_off = S.misc + 4;           // first entry of round r in rlist
    if (tid == 0) {
        uint32_t o = 0;
        for (int r = 0; r < kRounds; r++) { rl_off[r] = o; o += (uint32_t)(wmax / (r + 1) + 1); }
    }
    // lane -> slot class of the symbol binarisation for the slot-per-lane path (rangecoder::s, FFV1_RangeCoder.cpp:135-171):
    //   lane 0 zero flag | 1..10 exponent i = lane-1 (slot 10 also takes i > 9) | 11..21 sign for e = lane-11 |
    //   22..31 mantissa bit i = lane-22 (slot 31 also takes i > 9)
    const bool isB = lane >= 1 && lane <= 10, isD = lane >= 11 && lane <= 21;
    const int li = isB ? lane - 1 : isD ? lane - 11 : lane - 22;
    const int lslot = compact ? lane - (lane > 10) - 2 * (lane > 21) : lane;
    unsigned long long bins_total = 0;
#ifdef B200_PHASE_TIMING
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif

    for (int y = r0; y < r1; y++) {
        // software prefetch of the next payload row into registers; it lands in the ring after this row is coded
        int nx0[kMaxPixPerThread], nx1[kMaxPixPerThread];
        const bool have_next = y + 1 < r1;
        if (have_next) {
#pragma unroll
            for (int k = 0; k < kMaxPixPerThread; k++) {
                const int x = tid + k * kModelThreads;
                if (x < w) fetch(y + 1, x, nx0[k], nx1[k]);
            }
        }
        for (int pl = 0; pl < planes; pl++) {
            const int32_t* cur = S.ring + ((size_t)((y + 3) % 3) * planes + pl) * wmax;
            const int32_t* prv = S.ring + ((size_t)((y + 2) % 3) * planes + pl) * wmax;
            const int32_t* pp2 = S.ring + ((size_t)((y + 1) % 3) * planes + pl) * wmax;
            uint16_t* out = bins + pos;
            if (tid <= kRounds) S.rfill[tid] = 0;
            // ---- phase A (K2): prediction, context, fold, records per sample; per-chunk record totals and per-chunk counts of
            // every context class (class = ctx & 15 = owner warp)
            for (int k = 0; k < KC; k++) {
                const int c = warp * KC + k;
                const int x = c * 32 + lane;
                const bool valid = x < w;
                uint32_t nb = 0, cls = 32 + lane;
                if (valid) {
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
                    nb = d ? 2 * (31 - __clz(abs(d))) + 3 : 1;
                    cls = (uint32_t)ctx & (uint32_t)(NW - 1);
                }
                uint32_t incl = nb;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const uint32_t m = __match_any_sync(0xffffffffu, cls);
                if (c < kMaxChunks) {
                    if (lane < NW) S.ccnt[c * NW + lane] = 0;
                    if (lane == 31) S.ctot[c] = incl;
                }
                __syncwarp();
                if (valid) {
                    S.off[x] = incl - nb;                                      // chunk-relative for now
                    const uint32_t intra = __popc(m & lt);
                    S.tmp8[x] = (uint8_t)intra;
                    if (intra == 0) S.ccnt[c * NW + cls] = (uint16_t)__popc(m);
                }
            }
            __syncthreads();
            PHASE_MARK(0);
            // ---- phase B1: exclusive prefix over chunks of the class counts (warp q: class q) and of the record totals (warp 0)
            {
                const int q = warp;
                const int c0 = lane, c1 = lane + 32;
                uint32_t v0 = c0 < nchunk ? S.ccnt[c0 * NW + q] : 0u, v1 = c1 < nchunk ? S.ccnt[c1 * NW + q] : 0u;
                uint32_t i0 = v0, i1 = v1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                    if (lane >= o) { i0 += t0; i1 += t1; }
                }
                const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
                if (c0 < nchunk) S.ccnt[c0 * NW + q] = (uint16_t)(i0 - v0);
                if (c1 < nchunk) S.ccnt[c1 * NW + q] = (uint16_t)(tot0 + i1 - v1);
                if (lane == 0) S.clstot[q] = (uint16_t)(tot0 + tot1);
                if (warp == 0) {
                    uint32_t a0 = c0 < nchunk ? S.ctot[c0] : 0u, a1 = c1 < nchunk ? S.ctot[c1] : 0u;
                    uint32_t j0 = a0, j1 = a1;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t t0 = __shfl_up_sync(0xffffffffu, j0, o), t1 = __shfl_up_sync(0xffffffffu, j1, o);
                        if (lane >= o) { j0 += t0; j1 += t1; }
                    }
                    const uint32_t s0 = __shfl_sync(0xffffffffu, j0, 31), s1 = __shfl_sync(0xffffffffu, j1, 31);
                    if (c0 < nchunk) S.ctot[c0] = j0 - a0;
                    if (c1 < nchunk) S.ctot[c1] = s0 + j1 - a1;
                    if (lane == 0) S.misc[0] = s0 + s1;                        // records of this plane-row
                }
            }
            __syncthreads();
            PHASE_MARK(1);
            // ---- phase B2: final record offsets; samples partitioned by owner warp, x order kept
            const uint32_t total = S.misc[0];
            {
                uint32_t ct = lane < NW ? S.clstot[lane] : 0u, ci = ct;
#pragma unroll
                for (int o = 1; o < NW; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, ci, o); if (lane >= o) ci += t; }
                const uint32_t cst = ci - ct;                                   // lane q < 16: first plist entry of class q
                if (warp == 0 && lane < NW) S.cstart[lane] = (uint16_t)cst;
                for (int k = 0; k < KC; k++) {
                    const int c = warp * KC + k;
                    const int x = c * 32 + lane;
                    const bool valid = x < w;
                    const uint32_t cls = valid ? (S.ctx[x] & (uint32_t)(NW - 1)) : 0u;
                    const uint32_t base = __shfl_sync(0xffffffffu, cst, cls);
                    if (valid) {
                        S.off[x] += S.ctot[c];
                        S.plist[base + S.ccnt[c * NW + cls] + S.tmp8[x]] = (uint16_t)x;
                    }
                }
            }
            __syncthreads();
            PHASE_MARK(2);
            // ---- phase C: warp q ranks the samples of its context class: round = occurrences of the context earlier in the row
            {
                const uint32_t n = S.clstot[warp], base = S.cstart[warp];
                for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                    const bool valid = i0 + lane < n;
                    const uint32_t x = valid ? S.plist[base + i0 + lane] : 0u;
                    const uint32_t cx = valid ? S.ctx[x] : (0x10000u | lane);
                    const uint32_t m = __match_any_sync(0xffffffffu, cx);
                    const uint32_t b0 = valid ? S.cnt8[cx] : 0u;
                    const uint32_t r = b0 + __popc(m & lt);
                    __syncwarp();
                    if (valid && (m >> lane) == 1u) S.cnt8[cx] = (uint8_t)min(255u, b0 + __popc(m));   // highest lane of the group
                    __syncwarp();
                    const uint32_t rr = valid ? min(r, (uint32_t)kRounds) : (32u + lane);
                    const uint32_t m2 = __match_any_sync(0xffffffffu, rr);
                    const int lead = __ffs(m2) - 1;
                    uint32_t bs = 0;
                    if (valid && lead == lane) bs = atomicAdd(&S.rfill[rr], (uint32_t)__popc(m2));
                    bs = __shfl_sync(0xffffffffu, bs, lead);
                    if (valid) {
                        S.tmp8[x] = (uint8_t)rr;
                        if (rr < (uint32_t)kRounds) S.rlist[rl_off[rr] + bs + __popc(m2 & lt)] = (uint16_t)x;
                    }
                }
                __syncwarp();
                for (uint32_t i0 = lane; i0 < n; i0 += 32) S.cnt8[S.ctx[S.plist[base + i0]]] = 0;
            }
            __syncthreads();
            PHASE_MARK(3);
            // ---- phase D (K3): two samples per lane (independent chains, interleaved), all their bins; round by round (contexts
            // inside a round are distinct, so the order inside a round is free)
            for (int r = 0; r < kRounds; r++) {
                const uint32_t nr = S.rfill[r];
                if (nr == 0) break;
                for (uint32_t i0 = (uint32_t)warp * 64; i0 < nr; i0 += kModelThreads * 2) {
                    uint32_t nb[2] = {0, 0}, o[2] = {0, 0}, lo[2] = {0, 0}, hi[2] = {0, 0}, sbase[2] = {0, 0}, swz[2] = {0, 0};
                    uint32_t lw[2][9];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint32_t j = i0 + q * 32 + lane;
                        uint32_t e = 0;
                        if (j < nr) {
                            const uint32_t x = S.rlist[rl_off[r] + j];
                            const int v = S.val[x];
                            const uint32_t cxd = S.ctx[x];
                            o[q] = S.off[x];
                            sbase[q] = compact ? cxd * 27u : cxd * 32u;
                            swz[q] = compact ? 0u : ((cxd >> 2) & 7u) << 2;
                            const uint32_t a = (uint32_t)abs(v);
                            if (v == 0) { nb[q] = 1; lo[q] = 1; }
                            else {
                                e = 31 - __clz(a);
                                nb[q] = 2 * e + 3;
                                const uint32_t mant = e ? (__brev(a & ((1u << e) - 1u)) >> (32 - e)) : 0u;   // bit i of a -> bit e-1-i
                                const uint64_t bw = (uint64_t)(((1u << e) - 1u) << 1) | ((uint64_t)mant << (e + 2)) | ((uint64_t)(v < 0) << (2 * e + 2));
                                lo[q] = (uint32_t)bw; hi[q] = (uint32_t)(bw >> 32);
                            }
                        }
                        const uint32_t* lrow = reinterpret_cast<const uint32_t*>(S.slut + e * kSlutRow);
#pragma unroll
                        for (int k = 0; k < 9; k++) lw[q][k] = lrow[k];
                    }
                    const uint32_t nbmax = __reduce_max_sync(0xffffffffu, max(nb[0], nb[1]));
                    const bool fits = __all_sync(0xffffffffu, o[0] + nb[0] <= stage_cap && o[1] + nb[1] <= stage_cap);
                    uint8_t* sb0 = S.states + sbase[0];
                    uint8_t* sb1 = S.states + sbase[1];
                    if (fits) {
                        uint16_t* so0 = S.stage + o[0];
                        uint16_t* so1 = S.stage + o[1];
#pragma unroll
                        for (uint32_t b = 0; b < 35; b++) {
                            if (b >= nbmax) break;
                            const bool p0 = b < nb[0], p1 = b < nb[1];
                            uint8_t* s0 = sb0 + (((lw[0][b >> 2] >> ((b & 3) * 8)) & 255u) ^ swz[0]);
                            uint8_t* s1 = sb1 + (((lw[1][b >> 2] >> ((b & 3) * 8)) & 255u) ^ swz[1]);
                            uint32_t st0 = 0, st1 = 0;
                            if (p0) st0 = *s0;
                            if (p1) st1 = *s1;
                            const uint32_t bit0 = b < 32 ? (lo[0] >> b) & 1u : (hi[0] >> (b - 32)) & 1u;
                            const uint32_t bit1 = b < 32 ? (lo[1] >> b) & 1u : (hi[1] >> (b - 32)) & 1u;
                            const uint32_t t0 = S.t2[(bit0 << 8) | st0];
                            const uint32_t t1 = S.t2[(bit1 << 8) | st1];
                            if (p0) { *s0 = (uint8_t)t0; so0[b] = (uint16_t)(t0 >> 16); }
                            if (p1) { *s1 = (uint8_t)t1; so1[b] = (uint16_t)(t1 >> 16); }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            uint8_t* sbq = q ? sb1 : sb0;
                            for (uint32_t b = 0; b < nbmax; b++) {
                                if (b < nb[q]) {
                                    const uint32_t e = (nb[q] - 3) >> 1;
                                    uint8_t* sp = sbq + ((nb[q] == 1 ? 0u : (uint32_t)S.slut[e * kSlutRow + b]) ^ swz[q]);
                                    const uint32_t bit = b < 32 ? (lo[q] >> b) & 1u : (hi[q] >> (b - 32)) & 1u;
                                    const uint32_t tt = S.t2[(bit << 8) | *sp];
                                    *sp = (uint8_t)tt;
                                    const uint32_t idx = o[q] + b;
                                    if (idx < stage_cap) S.stage[idx] = (uint16_t)(tt >> 16); else out[idx] = (uint16_t)(tt >> 16);
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
            PHASE_MARK(4);
            // ---- phase E: samples beyond kRounds occurrences of their context (flat areas): the owner warp walks them in x order,
            // one sample per step, lane s = slot s of the context (slots are independent chains)
            if (S.rfill[kRounds]) {
                const uint32_t n = S.clstot[warp], base = S.cstart[warp];
                for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                    const bool valid = i0 + lane < n;
                    const uint32_t xl = valid ? S.plist[base + i0 + lane] : 0u;
                    uint32_t todo = __ballot_sync(0xffffffffu, valid && S.tmp8[xl] == kRounds);
                    while (todo) {
                        const int j = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t x = __shfl_sync(0xffffffffu, xl, j);
                        const int v = S.val[x];                 // warp-uniform
                        const uint32_t ob = S.off[x];
                        const uint32_t a = (uint32_t)abs(v);
                        const int e = 31 - __clz(a | 1);
                        uint8_t* sp = S.states + state_off(S.ctx[x], (uint32_t)lslot, compact);
                        if (e <= 8) {
                            const bool nz = v != 0;
                            const bool has = lane == 0 ? true : (nz && (isB ? li <= e : isD ? li == e : li < e));
                            const uint32_t bit = lane == 0 ? !nz : isB ? (uint32_t)(li < e) : isD ? (uint32_t)(v < 0) : ((a >> li) & 1u);
                            const int idx = lane == 0 ? 0 : isB ? 1 + li : isD ? 2 * e + 2 : 2 * e + 1 - li;
                            if (has) {
                                const uint32_t st = *sp;
                                const uint32_t rec = make_rec(st, bit);
                                if (ob + idx < stage_cap) S.stage[ob + idx] = (uint16_t)rec; else out[ob + idx] = (uint16_t)rec;
                                *sp = S.trans[(bit << 8) | st];
                            }
                        } else {
                            int n2 = 0, i0b = 0, step = 0;     // n2 bins; bin k uses index i = i0b + k*step
                            if (lane == 0) n2 = 1;
                            else if (lane <= 9) { n2 = 1; i0b = lane - 1; }
                            else if (lane == 10) { n2 = e - 8; i0b = 9; step = 1; }
                            else if (lane <= 21) { n2 = (lane - 11) == min(e, 10); }
                            else if (lane <= 30) { n2 = 1; i0b = lane - 22; }
                            else { n2 = e - 9; i0b = e - 1; step = -1; }
                            if (n2) {
                                uint32_t st = *sp;
                                for (int k = 0; k < n2; k++) {
                                    const int i = i0b + k * step;
                                    uint32_t bit, idx;
                                    if (lane == 0) { bit = 0; idx = 0; }
                                    else if (lane <= 10) { bit = i < e; idx = 1 + i; }
                                    else if (lane <= 21) { bit = v < 0; idx = 2 * e + 2; }
                                    else { bit = (a >> i) & 1; idx = 2 * e + 1 - i; }
                                    const uint32_t rec = make_rec(st, bit);
                                    if (ob + idx < stage_cap) S.stage[ob + idx] = (uint16_t)rec; else out[ob + idx] = (uint16_t)rec;
                                    st = S.trans[(bit << 8) | st];
                                }
                                *sp = (uint8_t)st;
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
            PHASE_MARK(5);
            // ---- phase F: staged records -> global, coalesced; pad the segment to a whole 128-byte block with no-op records
            {
                const uint32_t ns = min(total, stage_cap);
                if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
                    const uint32_t nv = ns >> 3;
                    const uint4* sv = reinterpret_cast<const uint4*>(S.stage);
                    uint4* dv = reinterpret_cast<uint4*>(out);
                    for (uint32_t i = tid; i < nv; i += kModelThreads) dv[i] = sv[i];
                    for (uint32_t i = (nv << 3) + tid; i < ns; i += kModelThreads) out[i] = S.stage[i];
                } else {
                    for (uint32_t i = tid; i < ns; i += kModelThreads) out[i] = S.stage[i];
                }
                PHASE_MARK(7);
                const uint32_t seg_len = seg_extra + total;
                const uint32_t pad = ((seg_len + 63u) & ~63u) - seg_len;
                if (tid < (int)pad) out[total + tid] = (uint16_t)kNopRec;
                if (tid == 0) A.rowcnt[(fs * A.band_rows + (y - r0)) * 3 + (ps ? 1 + pl : 0)] = seg_len;
                pos += total + pad;
                seg_extra = 0;
                bins_total += total;
            }
            __syncthreads();
            PHASE_MARK(6);
        }
        if (have_next) {
            int32_t* dst = S.ring + (size_t)((y + 4) % 3) * planes * wmax;
#pragma unroll
            for (int k = 0; k < kMaxPixPerThread; k++) {
                const int x = tid + k * kModelThreads;
                if (x < w) { dst[x] = nx0[k]; if (ps) dst[wmax + x] = nx1[k]; }
            }
            __syncthreads();
        }
        PHASE_MARK(7);
    }
#ifdef B200_PHASE_TIMING
    if (tid == 0) for (int k = 0; k < 8; k++) atomicAdd(reinterpret_cast<unsigned long long*>(A.flags + 16) + k, (unsigned long long)ph[k]);
#endif
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(A.flags + 2), bins_total);
    if (r1 < g.h) {   // carry the states to the next band
        const int n16 = state_bytes >> 4;
        const uint4* s = reinterpret_cast<const uint4*>(S.states);
        uint4* d = reinterpret_cast<uint4*>(save);
        for (int i = tid; i < n16; i += kModelThreads) d[i] = s[i];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The range coder, split in two.
//
// A binary range coder's output is one big number: every bin with value 1 adds (range - range1) at the byte position the
// coder has reached (FFV1_RangeCoder.cpp:71-102 seen from the encoder side: low += range - range1; on renormalisation the
// top byte of the 16-bit window leaves). Only `range` and the number of renormalisations form a serial recurrence.
//
//   k_range  one lane per (frame, slice): runs just that recurrence — range' = renorm((range * sp + c) >> 8) and the
//            shift count — over the slice's records, ~9 instructions per bin, branch-free, records staged through
//            shared memory with cp.async. Every 64 records (one 128-byte block) it leaves a checkpoint (range, bytes so far).
//   k_emit   one thread per 64-record block, fully parallel: replays the block from its checkpoint with the complete
//            coder step and adds the block's contribution into the slice's byte stream, held as big-endian 32-bit words,
//            with atomic adds (neighbouring blocks overlap by the 16-bit window and by carries; addition commutes).
//
// Slice header bins are ordinary records at the head of the Y stream (written by k_model), the terminator bin only moves
// `range` (k_range), and the coder's final flush is a single +0xFF (k_range). CRC and footer: k_pack.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// add v into big-endian word idx of the slice stream, propagating wrap-arounds towards the front
__device__ __forceinline__ void stream_add(uint32_t* W, int64_t idx, uint32_t v) {
    while (v && idx >= 0) {
        const uint32_t old = atomicAdd(W + idx, v);
        v = (old + v < old) ? 1u : 0u;
        idx--;
    }
}

__device__ __forceinline__ void range_step(uint32_t rec, uint32_t& range, uint32_t& cnt) {
    const uint32_t sp = rec & 0x1FFu;
    const uint32_t x = range * sp + ((rec & 0x200u) ? 0u : 255u);   // bit 0: range - ((range*state)>>8) == (range*(256-state)+255)>>8
    const bool sh = x < 0x10000u;
    range = sh ? (x & 0xFFFF00u) : (x >> 8);
    cnt += sh;
}

constexpr int kRingBlocks = 4;          // 128-byte blocks in flight per lane
constexpr int kMaxBandRows = 64;

__global__ void __launch_bounds__(32) k_range(const __grid_constant__ EncArgs A, int band, int nframes) {
    __shared__ uint4 s_ring[kRingBlocks][8][32];
    __shared__ uint16_t s_cnt[kMaxBandRows * 3][32];                // blocks per segment (plane-row)
    const int lane = threadIdx.x;
    const int gid = blockIdx.x * 32 + lane;
    const bool in_range = gid < nframes * A.nslices;
    const int slice = in_range ? gid % A.nslices : 0;
    const SliceGeom g = A.geom[slice];
    const int r0 = band * A.band_rows;
    const bool valid = in_range && r0 < g.h;
    const int r1 = min(r0 + A.band_rows, g.h);
    const int nseg = valid ? (r1 - r0) * 3 : 0;
    const size_t gsafe = in_range ? (size_t)gid : 0;

    uint32_t total = 0, usedY = 0, usedC = 0;                       // 64-record blocks this lane consumes in this band
    if (valid) {
        const uint32_t* rc = A.rowcnt + gsafe * A.band_rows * 3;
        for (int s = 0; s < nseg; s++) {
            const uint32_t c = (rc[s] + 63u) >> 6;
            s_cnt[s][lane] = (uint16_t)c;
            total += c;
            if (s % 3 == 0) usedY += c; else usedC += c;
        }
    }
    if (in_range) { A.used[gsafe * 2] = usedY; A.used[gsafe * 2 + 1] = usedC; }
    uint32_t maxb = total;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxb = max(maxb, __shfl_xor_sync(0xffffffffu, maxb, o));

    uint32_t range = 0xFF00u, cnt = 0;
    if (valid && band > 0) { const CoderState s = A.cstate[gid]; range = s.range; cnt = s.pos; }

    const uint4* pY = reinterpret_cast<const uint4*>(A.binsY + gsafe * A.capY);
    const uint4* pC = reinterpret_cast<const uint4*>(A.binsC + gsafe * A.capC);
    uint2* kY = A.ckptY + gsafe * (A.capY >> 6);
    uint2* kC = A.ckptC + gsafe * (A.capC >> 6);
    // two cursors over the lane's segments (Y row, Cb row, Cr row, Y row, ...): prefetch runs kRingBlocks-1 blocks ahead
    int pf_seg = -1, pf_pl = 2, cs_seg = -1, cs_pl = 2;
    uint32_t pf_rem = 0, cs_rem = 0;
    auto next_src = [&]() -> const uint4* {
        while (pf_rem == 0) {
            if (++pf_seg >= nseg) return nullptr;
            pf_pl = pf_pl == 2 ? 0 : pf_pl + 1;
            pf_rem = s_cnt[pf_seg][lane];
        }
        pf_rem--;
        const uint4* p = pf_pl == 0 ? pY : pC;
        if (pf_pl == 0) pY += 8; else pC += 8;
        return p;
    };
    auto next_ckpt = [&]() -> uint2* {
        while (cs_rem == 0) {
            if (++cs_seg >= nseg) return nullptr;
            cs_pl = cs_pl == 2 ? 0 : cs_pl + 1;
            cs_rem = s_cnt[cs_seg][lane];
        }
        cs_rem--;
        return cs_pl == 0 ? kY++ : kC++;
    };
    auto prefetch = [&](int slot) {
        const uint4* p = next_src();
        if (p) {
#pragma unroll
            for (int u = 0; u < 8; u++) cp_async16(&s_ring[slot][u][lane], p + u);
        }
        cp_async_commit();
    };
    for (int i = 0; i < kRingBlocks - 1; i++) prefetch(i);
    const uint32_t nop2 = kNopRec | (kNopRec << 16);
    for (uint32_t bi = 0; bi < maxb; bi++) {
        prefetch((bi + kRingBlocks - 1) % kRingBlocks);
        cp_async_wait<kRingBlocks - 1>();
        const bool live = bi < total;
        uint2* ck = next_ckpt();
        if (ck) *ck = make_uint2(range, cnt);
        const int slot = bi % kRingBlocks;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 q = s_ring[slot][u][lane];
            if (!live) q = make_uint4(nop2, nop2, nop2, nop2);
            range_step(q.x & 0xFFFFu, range, cnt); range_step(q.x >> 16, range, cnt);
            range_step(q.y & 0xFFFFu, range, cnt); range_step(q.y >> 16, range, cnt);
            range_step(q.z & 0xFFFFu, range, cnt); range_step(q.z >> 16, range, cnt);
            range_step(q.w & 0xFFFFu, range, cnt); range_step(q.w >> 16, range, cnt);
        }
    }
    cp_async_wait<0>();
    if (valid) {
        if (r1 == g.h) {
            range_step(127u, range, cnt);                          // terminator: state 129, bit 0 (FFV1_Slice.cpp:334-343)
            // final flush of the coder (so that the decoder's BytesUsed() lands on the end, FFV1_Slice.cpp:297-299):
            // low += 0xFF in the current 16-bit window (stream bytes cnt, cnt+1), then two forced renormalisations
            uint32_t* W = reinterpret_cast<uint32_t*>(A.scratch + gsafe * A.slice_cap);
            const uint32_t q = cnt + 1;
            if ((size_t)q + 8 < A.slice_cap) stream_add(W, q >> 2, 0xFFu << ((3 - (q & 3)) * 8));
            else atomicOr(A.flags, 1u);
            A.slice_size[gid] = cnt + 1;                           // cnt + 2 bytes produced; the last stays inside the coder
        } else {
            CoderState s;
            s.low = 0; s.range = range; s.pending = 0; s.run = 0; s.pos = cnt; s.crc = 0; s.offY = 0; s.offC = 0;
            A.cstate[gid] = s;
        }
    }
}

// k_emit: grid (blocks of 256 coder-blocks, stream Y/C, frame*slice)
constexpr int kEmitThreads = 256;
constexpr int kEmitWords = 22;          // local window: word 0 spare, then stream words (c0>>2)-1 ... (c0>>2)+19

__global__ void __launch_bounds__(kEmitThreads) k_emit(const __grid_constant__ EncArgs A) {
    __shared__ uint32_t s_loc[kEmitWords][kEmitThreads];
    const int gid = blockIdx.z, stream = blockIdx.y, tid = threadIdx.x;
    const uint32_t used = A.used[(size_t)gid * 2 + stream];
    if (blockIdx.x * kEmitThreads >= used) return;
    const uint32_t blk = blockIdx.x * kEmitThreads + tid;
    if (blk >= used) return;
    const size_t cap = stream ? A.capC : A.capY;
    const uint4* src = reinterpret_cast<const uint4*>((stream ? A.binsC : A.binsY) + (size_t)gid * cap) + (size_t)blk * 8;
    const uint2 ck = (stream ? A.ckptC : A.ckptY)[(size_t)gid * (cap >> 6) + blk];
    uint4 q[8];
#pragma unroll
    for (int u = 0; u < 8; u++) q[u] = __ldg(src + u);
#pragma unroll
    for (int i = 0; i < kEmitWords; i++) s_loc[i][tid] = 0;

    uint32_t range = ck.x;
    const uint32_t c0 = ck.y;
    // byte `lb` of the local window is stream byte 4*((c0>>2)-1) + lb - 4; the block starts at local byte 8 + (c0&3)
    uint32_t lpos = 8 + (c0 & 3);           // local byte index of the coder window's high byte
    uint64_t acc = 0;                       // bits 0..15 window, then pending bytes, then carry
    uint32_t nsh = 0;
    auto flush = [&](uint32_t extra) {
        // the value above the window (pending bytes + carry) has its LSB at the end of local byte lpos + nsh/8 - 1;
        // `extra` = 16 also retires the window itself (end of block)
        const uint32_t k = nsh >> 3;
        const uint64_t V = extra ? acc : (acc >> 16);
        const uint32_t lb = lpos + k - 1 + (extra >> 3);
        const uint32_t wi = lb >> 2, sft = (3 - (lb & 3)) * 8;
        const uint64_t add = V << sft;
        uint64_t cur = ((uint64_t)s_loc[wi - 1][tid] << 32) | s_loc[wi][tid];
        const uint64_t sum = cur + add;
        s_loc[wi - 1][tid] = (uint32_t)(sum >> 32);
        s_loc[wi][tid] = (uint32_t)sum;
        if (sum < cur) { int j = (int)wi - 2; while (j >= 0 && ++s_loc[j][tid] == 0) j--; }
        lpos += k;
        acc &= 0xFFFFull;
        nsh = 0;
    };
    auto step = [&](uint32_t rec) {
        const uint32_t sp = rec & 0x1FFu;
        const uint32_t bit = (rec >> 9) & 1u;
        const uint32_t x = range * sp + (bit ? 0u : 255u);
        const uint32_t nr = x >> 8;
        acc += bit ? (uint64_t)(range - nr) : 0ull;
        const uint32_t sh = x < 0x10000u ? 8u : 0u;
        range = nr << sh;
        acc <<= sh;
        nsh += sh;
    };
#pragma unroll
    for (int u = 0; u < 8; u++) {
        step(q[u].x & 0xFFFFu); step(q[u].x >> 16); step(q[u].y & 0xFFFFu); step(q[u].y >> 16);
        flush(0);
        step(q[u].z & 0xFFFFu); step(q[u].z >> 16); step(q[u].w & 0xFFFFu); step(q[u].w >> 16);
        flush(0);
    }
    flush(16);
    uint32_t* W = reinterpret_cast<uint32_t*>(A.scratch + (size_t)gid * A.slice_cap);
    const int64_t w0 = (int64_t)(c0 >> 2) - 2;                      // stream word of local word 0
    if ((size_t)(c0 + 80) >= A.slice_cap) { atomicOr(A.flags, 1u); return; }
#pragma unroll 1
    for (int i = kEmitWords - 1; i >= 1; i--) {
        const uint32_t v = s_loc[i][tid];
        if (v) stream_add(W, w0 + i, v);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_scan: slice payload sizes -> slice sizes incl. footer -> arena offsets (single CTA; the arrays are tiny)
__global__ void __launch_bounds__(1024) k_scan(const __grid_constant__ EncArgs A, int nframes) {
    __shared__ uint64_t s_part[1024];
    const int n = nframes * A.nslices;
    const uint32_t foot = A.ec ? 8u : 3u;
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(b + per, n);
    uint64_t sum = 0;
    for (int i = b; i < e; i++) sum += A.slice_size[i] + foot;
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
        off += A.slice_size[i] + foot;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nframes; f += 1024) {
        const int first = f * A.nslices, last = first + A.nslices - 1;
        A.frame_off[f] = A.slice_off[first];
        A.frame_len[f] = A.slice_off[last] + A.slice_size[last] + foot - A.slice_off[first];
    }
    if (threadIdx.x == 1023 && s_part[1023] > A.arena_cap) atomicOr(A.flags, 2u);
}

// k_pack: one CTA per slice. Copies the payload from the slice scratch (big-endian 32-bit words, see k_emit) to its place in the packet arena, computes the
// slice CRC in parallel and appends the footer: slice_size (24 bit BE), error_status 0, crc parity (32 bit BE)
// (FFV1_Slice.cpp:247-253, :301-315). CRC = poly 0x04C11DB7, MSB first, init 0, no final xor — linear, so the CRC of a
// concatenation is crc(A) * x^(8|B|) + crc(B) in GF(2)[x]/P (ZenCRC32 semantics, ZenCRC32.cpp:1097-1135).
constexpr int kPackThreads = 256;
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b) {
    uint32_t r = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        r = (r << 1) ^ ((r >> 31) ? 0x04C11DB7u : 0u);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
__global__ void __launch_bounds__(kPackThreads) k_pack(const __grid_constant__ EncArgs A) {
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_red[kPackThreads / 32];
    const int gid = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 256; i += kPackThreads) s_tab[i] = A.crc_table[i];
    const uint32_t n = A.slice_size[gid];
    const uint64_t off = A.slice_off[gid];
    const uint32_t foot = A.ec ? 8u : 3u;
    if (off + n + foot > A.arena_cap || n + foot > A.slice_cap) return;
    const uint8_t* src = A.scratch + (size_t)gid * A.slice_cap;
    uint8_t* dst = A.arena + off;
    __syncthreads();
    // ---- CRC of the payload: thread t owns bytes [t*L, min((t+1)*L, n)), L a multiple of 16
    uint32_t crc = 0;
    if (A.ec) {
        const uint32_t L = (((n + kPackThreads - 1) / kPackThreads) + 15u) & ~15u;
        const uint32_t b0 = min(n, tid * L), b1 = min(n, b0 + L);
        uint32_t i = b0;
        for (; i + 16 <= b1; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + i);
            const uint32_t wd[4] = {bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w)};
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int bb = 0; bb < 4; bb++) crc = (crc << 8) ^ s_tab[(crc >> 24) ^ ((wd[k] >> (8 * bb)) & 255u)];
            }
        }
        for (; i < b1; i++) crc = (crc << 8) ^ s_tab[(crc >> 24) ^ src[i ^ 3]];
        // multiply by x^(8 * bytes that follow this thread's piece) with square-and-multiply on x^8
        uint32_t after = n - b1;
        uint32_t pw = 0x00000100u;                 // x^8
        while (after) {
            if (after & 1u) crc = gf_mulmod(crc, pw);
            after >>= 1;
            if (after) pw = gf_mulmod(pw, pw);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = crc;
    }
    // ---- copy (16-byte stores, funnel-shifted loads: src is 16-byte aligned, dst is not)
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    if (tid < (int)head) dst[tid] = src[tid ^ 3];
    const uint32_t body = (n - head) >> 4;
    const uint32_t sh = (head & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = tid; i < body; i += kPackThreads) {
        const uint32_t* p = sw + i * 4;
        uint32_t w0 = bswap32(p[0]), w1 = bswap32(p[1]), w2 = bswap32(p[2]), w3 = bswap32(p[3]), w4 = sh ? bswap32(p[4]) : 0;
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
        dv[i] = v;
    }
    const uint32_t tail0 = head + (body << 4);
    if (tid < (int)(n - tail0)) dst[tail0 + tid] = src[(tail0 + tid) ^ 3];
    __syncthreads();
    if (tid == 0) {
        uint8_t f[8] = {(uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n, 0, 0, 0, 0, 0};
        if (A.ec) {
            uint32_t c = 0;
            for (int i = 0; i < kPackThreads / 32; i++) c ^= s_red[i];
            for (int i = 0; i < 4; i++) c = (c << 8) ^ s_tab[(c >> 24) ^ f[i]];
            f[4] = (uint8_t)(c >> 24); f[5] = (uint8_t)(c >> 16); f[6] = (uint8_t)(c >> 8); f[7] = (uint8_t)c;
        }
        for (uint32_t i = 0; i < foot; i++) dst[n + i] = f[i];
    }
}

// ------------------------------------------------------------------------------------------------------------------
cudaError_t configure_kernels(int nctx, int sstride, int wmax, int stage_cap) {
    size_t need = model_smem_bytes(nctx, sstride, wmax, 2, stage_cap);
    return cudaFuncSetAttribute(k_model, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
}

cudaError_t launch_model(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    dim3 grid(a.nslices * 2, nframes);
    size_t smem = model_smem_bytes(a.nctx, a.sstride, a.wmax, 2, a.stage_cap);
    k_model<<<grid, kModelThreads, smem, s>>>(a, band);
    return cudaGetLastError();
}

cudaError_t launch_range(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    int n = nframes * a.nslices;
    k_range<<<(n + 31) / 32, 32, 0, s>>>(a, band, nframes);
    return cudaGetLastError();
}

cudaError_t launch_emit(const EncArgs& a, int nframes, cudaStream_t s) {
    int n = nframes * a.nslices;
    const unsigned nblk = (unsigned)((a.capC >> 6) + kEmitThreads - 1) / kEmitThreads;
    k_emit<<<dim3(nblk, 2, n), kEmitThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pack(const EncArgs& a, int nframes, cudaStream_t s) {
    k_scan<<<1, 1024, 0, s>>>(a, nframes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_pack<<<nframes * a.nslices, kPackThreads, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace b200
