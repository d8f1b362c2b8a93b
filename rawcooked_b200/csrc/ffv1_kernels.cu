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

#include <cstdlib>
#include <type_traits>

#include "../../include/b200enc.h"
#include "pixel_layouts.cuh"

namespace b200 {

__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

// ------------------------------------------------------------------------------------------------------------------
// Records. k_model hands every range-coder bin to k_range / k_emit as (q, bit): q = sp - 1 where
// sp = bit ? state : 256 - state is the 8-bit probability of the coded value (1..255), bit the coded value. q = 255
// (sp = 256, bit 0) leaves the coder untouched; it pads every plane-row segment to whole 128-record blocks. Inside
// k_model a record is the 16-bit value q | bit << 8 = 255 + (bit ? state : -state); in global memory the q bytes and the
// bits (one bit-plane, record r = bit r & 7 of byte r >> 3) are separate arrays.
constexpr uint32_t kNopRec = 0x00FFu;

// ------------------------------------------------------------------------------------------------------------------
// k_model
//
// One CTA per (frame, slice, plane-set) and band. Shared memory (dynamic):
//   states  nctx*sstride B   adaptive state of every (context, slot) of this plane-set (32 B per context, 28 for 8-bit streams)
//   raw     the payload bytes of the next slice row, as stored in the file (bulk copy); mbar: two mbarriers (states, row)
//   ring    3 rows x planes x wmax int32   RCT'd samples of rows y, y-1, y-2
//   qtab    5 x 256 int16    quantisation tables
//   tpow    one_state iterated 2^i times (runs of zeros); trans: next state by (bit, state)
//   ent     the samples of the plane-row being coded, grouped by context class (= owner warp), x order inside a class, 16 bytes
//           each: slots the symbol uses (bit mask), value of the bin of each slot (bit mask), byte offsets of its records in the
//           stage (2o | (2o + 4e) << 16), folded residual (after the context-sign flip, 18 bits) | context << 18
//   segcnt  [segment][class] samples of the class in a column segment (only when a plane-row needs more than one segment)
//   ctot    records per 32-sample chunk
//   wcnt    [class][warp] samples of the class among the chunks that warp prepared in S1
//   stage   the records of the plane-row (segment), 16 bit each, split into q bytes + bit-plane on the way out
//
// Per plane-row: S1 all samples in parallel: neighbours, context, residual, bin count, class masks -> S2 (K3) warp q
// resolves the adaptive states of the samples of class q, in bitstream order, 32 at a time: samples of one batch that use
// distinct contexts go one per lane with the 32 state bytes of the context in registers (all bins of the symbol, the steps
// of different slots overlap); repeated contexts wait for the next round of the batch; rounds with a few samples run one
// sample per step with one lane per slot -> S3 records to global memory. Warps never share a context, so there is no
// barrier inside S2.
struct ModelSmem {
    uint8_t* states; uint8_t* raw; uint64_t* mbar; int32_t* ring; int16_t* qtab;
    uint4* ent; uint32_t* ctot; uint16_t* wcnt; uint32_t* segcnt; uint32_t* misc; uint8_t* tpow; uint8_t* trans; uint16_t* stage;
};
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
// the bytes of one slice row of the payload, widened to 16-byte boundaries at both ends (6 bytes per pixel at most)
__host__ __device__ inline size_t model_raw_cap(int wmax) { return align16((size_t)wmax * 6 + 48); }

// bytes of everything except the record staging area
size_t model_smem_fixed(int nctx, int sstride, int wmax, int planes) {
    size_t n = align16((size_t)nctx * sstride);
    n += model_raw_cap(wmax) + 16;           // raw payload row + two mbarriers
    n += align16((size_t)3 * planes * wmax * 4);
    n += 5 * 256 * 2;
    n += (size_t)wmax * 16;                  // ent
    const int nch = (wmax + 31) / 32;
    n += align16((size_t)nch * 4);           // ctot
    n += align16((size_t)kModelWarps * kModelWarps * 2);      // wcnt
    n += (size_t)kMaxSeg * kModelWarps * 4;  // segcnt
    n += kModelWarps * 64;                   // misc: per-lane landing place of the records of idle lanes
    n += 5 * 256;                            // tpow
    n += 512;                                // trans
    return n;
}
size_t model_smem_bytes(int nctx, int sstride, int wmax, int planes, int stage_cap) {
    return model_smem_fixed(nctx, sstride, wmax, planes) + align16((size_t)stage_cap * 2);
}

__device__ __forceinline__ ModelSmem carve(uint8_t* base, int nctx, int sstride, int wmax, int planes) {
    ModelSmem m;
    m.states = base; base += align16((size_t)nctx * sstride);
    m.raw = base; base += model_raw_cap(wmax);
    m.mbar = reinterpret_cast<uint64_t*>(base); base += 16;
    m.ring = reinterpret_cast<int32_t*>(base); base += align16((size_t)3 * planes * wmax * 4);
    m.qtab = reinterpret_cast<int16_t*>(base); base += 5 * 256 * 2;
    m.ent = reinterpret_cast<uint4*>(base); base += (size_t)wmax * 16;
    const int nch = (wmax + 31) / 32;
    m.ctot = reinterpret_cast<uint32_t*>(base); base += align16((size_t)nch * 4);
    m.wcnt = reinterpret_cast<uint16_t*>(base); base += align16((size_t)kModelWarps * kModelWarps * 2);
    m.segcnt = reinterpret_cast<uint32_t*>(base); base += (size_t)kMaxSeg * kModelWarps * 4;
    m.misc = reinterpret_cast<uint32_t*>(base); base += kModelWarps * 64;
    m.tpow = base; base += 5 * 256;
    m.trans = base; base += 512;
    m.stage = reinterpret_cast<uint16_t*>(base);
    return m;
}


// -DB200_PHASE_TIMING: thread 0 of every CTA accumulates the cycles between phase boundaries into flags[16 + 2*phase]
#ifdef B200_PHASE_TIMING
#define PHASE_MARK(k) do { if (tid == 0) { long long t_ = clock64(); ph[k] += t_ - tlast; tlast = t_; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

// position of a slot inside the state row of a context: 8-bit streams (folded residual of 9 bits, exponent <= 8) never
// use slots 10, 20, 21, 30, 31, so their rows hold 27 states in 28 bytes
template <bool kCompact>
__host__ __device__ constexpr int cslot(int slot) { return kCompact ? slot - (slot > 10 ? 1 : 0) - (slot > 21 ? 2 : 0) : slot; }

// shared-memory accesses by 32-bit shared address (no generic-address conversion per access)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8_volatile(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }


// ---- bulk asynchronous copies (the 1-D form of TMA, UBLKCP in SASS) and the mbarrier they report to: the context-state table
// of an item (162 KB) comes in and goes out as one copy each, the payload row of the next slice row is in flight while this one
// is coded; no thread spends issue slots on either
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A symbol with exponent > 9 on the one-lane-per-slot path: slots 10 and 31 take several bins of the symbol; lane s runs the bins
// of its slot on the state it holds and returns the new state (rangecoder::s, FFV1_RangeCoder.cpp:135-171). Rare (|residual| >=
// 1024), kept out of line so that the sample loop stays small.
__device__ __noinline__ uint32_t wide_symbol(const uint4 E, uint32_t st, int lane, uint32_t stage_a, uint32_t trans_a) {
    const int vj = ((int)(E.w << 14)) >> 14;
    const uint32_t ob = (E.z & 0xFFFFu) >> 1;
    const uint32_t aj = (uint32_t)abs(vj);
    const int ej = 31 - __clz(aj | 1);
    int n2 = 0, i0b = 0, step = 0;     // n2 bins; bin k uses index i = i0b + k*step
    if (lane == 0) n2 = 1;
    else if (lane <= 9) { n2 = 1; i0b = lane - 1; }
    else if (lane == 10) { n2 = ej - 8; i0b = 9; step = 1; }
    else if (lane <= 21) { n2 = (lane - 11) == min(ej, 10); }
    else if (lane <= 30) { n2 = 1; i0b = lane - 22; }
    else { n2 = ej - 9; i0b = ej - 1; step = -1; }
    for (int kk = 0; kk < n2; kk++) {
        const int i = i0b + kk * step;
        bool bit; uint32_t bidx;
        if (lane == 0) { bit = false; bidx = 0; }
        else if (lane <= 10) { bit = i < ej; bidx = 1 + i; }
        else if (lane <= 21) { bit = vj < 0; bidx = 2 * ej + 2; }
        else { bit = ((aj >> i) & 1u) != 0; bidx = 2 * ej + 1 - i; }
        const uint32_t rec = bit ? 255u + st : 255u - st;
        sts_u16(stage_a + (ob + bidx) * 2u, rec);
        st = lds_u8(trans_a + st + (bit ? 256u : 0u));
    }
    return st;
}

template <bool kCompact>
__device__ __forceinline__ void k_model_body(const EncArgs& A, int band, int nframes, uint8_t* smem_raw, int slice, int ps, int frame, uint32_t& ph_state, uint32_t& ph_row) {
    const SliceGeom g = A.geom[slice];
    const int r0 = band * A.band_rows;
    if (r0 >= g.h) return;
    const int r1 = min(r0 + A.band_rows, g.h);
    const int planes = ps ? 2 : 1;
    const int w = g.w, wmax = A.wmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kModelThreads / 32;
    constexpr int kRow = kCompact ? 28 : 32;                   // state bytes per context
    const ModelSmem S = carve(smem_raw, A.nctx, kRow, wmax, 2);      // one layout for both plane-sets: the constant tables are loaded once
    const uint32_t stage_cap = (uint32_t)A.stage_cap;
    const uint32_t stage_a = smem_addr(S.stage), states_a = smem_addr(S.states), tpow_a = smem_addr(S.tpow), trans_a = smem_addr(S.trans);
    const uint32_t lt = (1u << lane) - 1u;

    const size_t fs = (size_t)frame * A.nslices + slice;
    const int state_bytes = (int)align16((size_t)A.nctx * kRow);
    uint8_t* save = A.state_save + (fs * 2 + ps) * (size_t)state_bytes;
    const uint32_t mb_state = smem_addr(S.mbar), mb_row = mb_state + 8u, raw_a = smem_addr(S.raw);
    // context states: 128 at the start of every frame (intra-only), else carried from the previous band: one bulk copy, waited
    // for after the first rows have been loaded (the constant tables were loaded once by k_model_loop)
    if (band == 0) {
        const int n16 = state_bytes >> 4;
        uint4* d = reinterpret_cast<uint4*>(S.states);
        const uint4 v = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        for (int i = tid; i < n16; i += kModelThreads) d[i] = v;
    } else if (tid == 0) {
        mbar_expect_tx(mb_state, (uint32_t)state_bytes);
        bulk_g2s(states_a, save, (uint32_t)state_bytes, mb_state);
    }
    const uint8_t* fin = A.in + (size_t)frame * A.frame_bytes;
    const int off = 1 << A.bits;

    // forward RCT of pixel x of a slice row whose payload row starts at `row` (inverse of Transform.cpp:29-37); p0 = Y or Cb, p1 = Cr
    auto fetch_from = [&](const uint8_t* row, int x, int& p0, int& p1) {
        int r, gg, b;
        load_rgb(row, A.layout, g.x0 + x, r, gg, b);
        if (A.swap_bg) { int t = gg; gg = b; b = t; }
        b -= gg; r -= gg; gg += (b + r) >> 2; b += off; r += off;
        if (ps == 0) { p0 = gg; p1 = 0; } else { p0 = b; p1 = r; }
    };
    auto fetch = [&](int y, int x, int& p0, int& p1) { fetch_from(fin + (size_t)(g.y0 + y) * A.row_bytes, x, p0, p1); };
    // the bytes of a payload row this slice reads: [first_off, end_off) from the start of the row
    uint32_t first_off, end_off;
    {
        const uint32_t x0 = (uint32_t)g.x0, x1 = (uint32_t)(g.x0 + w);
        switch (A.layout) {
            case B200_DPX_RGB_8: case B200_TIFF_RGB_8: first_off = 3u * x0; end_off = 3u * x1; break;
            case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: first_off = 4u * x0; end_off = 4u * x1; break;
            case B200_DPX_RGB_12_PACKED_BE: first_off = ((36u * x0) >> 5) * 4u; end_off = (((36u * x1 - 1u) >> 5) + 1u) * 4u; break;
            default: first_off = 6u * x0; end_off = 6u * x1; break;
        }
        if (end_off > A.row_bytes) end_off = A.row_bytes;
    }
    const uintptr_t in_end = (reinterpret_cast<uintptr_t>(A.in) + (size_t)nframes * A.frame_bytes) & ~(uintptr_t)15;
    // slice row y as a bulk copy: the 16-byte aligned window around its bytes (none when the window would pass the end of the input)
    auto row_window = [&](int y, uintptr_t& wlo, uint32_t& len) -> bool {
        const uintptr_t R = reinterpret_cast<uintptr_t>(fin + (size_t)(g.y0 + y) * A.row_bytes);
        wlo = (R + first_off) & ~(uintptr_t)15;
        const uintptr_t whi = (R + end_off + 15) & ~(uintptr_t)15;
        len = (uint32_t)(whi - wlo);
        return whi <= in_end;
    };
    auto issue_row = [&](int y) {          // one thread
        uintptr_t wlo; uint32_t len;
        if (row_window(y, wlo, len)) {
            mbar_expect_tx(mb_row, len);
            bulk_g2s(raw_a, reinterpret_cast<const void*>(wlo), len, mb_row);
        }
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

    const size_t cap = ps ? A.capC : A.capY;
    uint8_t* gq = (ps ? A.qC : A.qY) + fs * cap;
    uint8_t* gb = (ps ? A.bC : A.bY) + fs * (cap >> 3);
    uint32_t pos = 0;                       // records emitted so far in this band (CTA-uniform, multiple of 128)
    uint32_t seg_extra = 0;                 // records already staged ahead of the next plane-row (the slice header)
    if (band == 0 && ps == 0) {
        const int nh = A.hdr_cnt[slice];
        for (int i = tid; i < nh; i += kModelThreads) S.stage[i] = A.hdr_bins[(size_t)slice * kMaxHeaderBins + i];
        seg_extra = (uint32_t)nh;
    }
    if (tid == 0 && r0 + 1 < r1) issue_row(r0 + 1);
    if (band != 0) { mbar_wait(mb_state, ph_state); ph_state ^= 1u; }
    __syncthreads();

    const int nchunk = (w + 31) >> 5;
    const int KC = (nchunk + NW - 1) / NW;  // 32-sample chunks per warp in S1: chunk c belongs to warp c % NW
    const int sbits = A.sbits;
    // one-lane-per-slot path: lane -> slot class of the symbol binarisation (rangecoder::s, FFV1_RangeCoder.cpp:135-171):
    //   lane 0 zero flag | 1..10 exponent i = lane-1 (slot 10 also takes i > 9) | 11..21 sign for e = lane-11 |
    //   22..31 mantissa bit i = lane-22 (slot 31 also takes i > 9)
    const bool isB = lane >= 1 && lane <= 10, isD = lane >= 11 && lane <= 21;
    const bool lane_has_slot = !kCompact || !(lane == 10 || lane == 20 || lane == 21 || lane >= 30);
    const int lslot = cslot<kCompact>(lane);
    unsigned long long bins_total = 0;
#ifdef B200_PHASE_TIMING
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif

    for (int y = r0; y < r1; y++) {
        const bool have_next = y + 1 < r1;
        for (int pl = 0; pl < planes; pl++) {
            const int32_t* cur = S.ring + ((size_t)((y + 3) % 3) * planes + pl) * wmax;
            const int32_t* prv = S.ring + ((size_t)((y + 2) % 3) * planes + pl) * wmax;
            const int32_t* pp2 = S.ring + ((size_t)((y + 1) % 3) * planes + pl) * wmax;
            // ---- S1 (K2): prediction, context, fold, records per sample. Warp w prepares chunks [w KC, (w+1) KC) and keeps
            // them in registers until the class lists can be laid out (after the barrier)
            constexpr int KCMAX = (2048 / 32 + NW - 1) / NW;                       // wmax <= 2048
            if (lane < NW) S.wcnt[lane * NW + warp] = 0;
            if (tid < kMaxSeg * NW) S.segcnt[tid] = 0;
            __syncwarp();
            int sv[KCMAX];
            uint32_t scx[KCMAX], sof[KCMAX], spi[KCMAX];   // context | class << 16 (all ones: no sample), first record within the chunk, rank among my warp's samples of the class
#pragma unroll
            for (int k = 0; k < KCMAX; k++) {
                const int c = warp * KC + k;
                sv[k] = 0; scx[k] = 0xFFFFFFFFu; sof[k] = 0; spi[k] = 0;
                if (k < KC && c < nchunk) {
                    const int x = c * 32 + lane;
                    const bool valid = x < w;
                    uint32_t nb = 0, cls = 32u + (uint32_t)lane;
                    int d = 0, ctx = 0;
                    if (valid) {
                        const int T = prv[x];
                        const int RT = prv[x + 1 < w ? x + 1 : w - 1];            // sample[1][w] = sample[1][w-1]
                        const int L = x > 0 ? cur[x - 1] : prv[0];                 // sample[0][-1] = sample[1][0]
                        const int LT = x > 0 ? prv[x - 1] : pp2[0];                // what sample[1][-1] was set to one row earlier
                        ctx = S.qtab[(L - LT) & 255] + S.qtab[256 + ((LT - T) & 255)] + S.qtab[512 + ((T - RT) & 255)];
                        if (A.is5) {
                            const int LL = x > 1 ? cur[x - 2] : (x == 1 ? prv[0] : 0);
                            const int TT = pp2[x];
                            ctx += S.qtab[768 + ((LL - L) & 255)] + S.qtab[1024 + ((TT - T) & 255)];
                        }
                        d = cur[x] - median3(L, L + T - LT, T);
                        if (ctx < 0) { ctx = -ctx; d = -d; }
                        d = (d << (32 - sbits)) >> (32 - sbits);                  // fold to sbits, sign-extended
                        cls = NW == 16 ? ((uint32_t)ctx * 2654435761u) >> 28 : NW == 32 ? ((uint32_t)ctx * 2654435761u) >> 27
                                                     : ((((uint32_t)ctx * 2654435761u) >> 16) * (uint32_t)NW) >> 16;   // multiplicative hash: even class sizes
                        nb = d ? 2 * (31 - __clz(abs(d))) + 3 : 1;
                    }
                    uint32_t incl = nb;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    const uint32_t m = __match_any_sync(0xffffffffu, cls);
                    const int leader = __ffs(m) - 1;
                    uint32_t old = 0;
                    if (valid && lane == leader) {                                  // my warp's running count of the class
                        old = S.wcnt[cls * NW + warp];
                        S.wcnt[cls * NW + warp] = (uint16_t)(old + (uint32_t)__popc(m));
                    }
                    old = __shfl_sync(0xffffffffu, old, leader);
                    __syncwarp();
                    if (valid) { sv[k] = d; scx[k] = (uint32_t)ctx | (cls << 16); sof[k] = incl - nb; spi[k] = old + (uint32_t)__popc(m & lt); }
                    if (lane == 31) S.ctot[c] = incl;
                }
            }
            __syncthreads();
            PHASE_MARK(0);
            // the next payload row (in flight since the previous row, or since the prologue) may now replace row y-2 in the ring
            // (nobody reads that one any more)
            bool row_taken = false;
            if (have_next && pl == planes - 1) {
                int32_t* dst = S.ring + (size_t)((y + 4) % 3) * planes * wmax;
                uintptr_t wlo; uint32_t len;
                const uint8_t* rowp;
                if (row_window(y + 1, wlo, len)) {
                    mbar_wait(mb_row, ph_row);
                    ph_row ^= 1u;
                    rowp = S.raw + (ptrdiff_t)(reinterpret_cast<uintptr_t>(fin + (size_t)(g.y0 + y + 1) * A.row_bytes) - wlo);
                } else {
                    rowp = fin + (size_t)(g.y0 + y + 1) * A.row_bytes;
                }
                for (int x = tid; x < w; x += kModelThreads) {
                    int p0, p1;
                    fetch_from(rowp, x, p0, p1);
                    dst[x] = p0;
                    if (ps) dst[wmax + x] = p1;
                }
                row_taken = true;
            }
            // exclusive prefix of the chunk totals, every warp for itself: lane l holds chunks l and l + 32
            uint32_t ex0, ex1, total;
            {
                const uint32_t t0 = lane < nchunk ? S.ctot[lane] : 0u, t1v = lane + 32 < nchunk ? S.ctot[lane + 32] : 0u;
                uint32_t i0 = t0, i1 = t1v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a0 = __shfl_up_sync(0xffffffffu, i0, o), a1 = __shfl_up_sync(0xffffffffu, i1, o);
                    if (lane >= o) { i0 += a0; i1 += a1; }
                }
                const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
                ex0 = i0 - t0;
                ex1 = tot0 + i1 - t1v;
                total = tot0 + __shfl_sync(0xffffffffu, i1, 31);
            }
            auto chunk_base = [&](int c) -> uint32_t {     // c warp-uniform; c == nchunk gives the total
                if (c >= nchunk) return total;
                return c < 32 ? __shfl_sync(0xffffffffu, ex0, c) : __shfl_sync(0xffffffffu, ex1, c - 32);
            };
            // column segments: chunks [sc[s], sc[s+1]) whose records fit the stage together
            int sc[kMaxSeg + 1];
            int nsg = 0;
            sc[0] = 0;
            if (seg_extra + total <= stage_cap) { sc[1] = nchunk; nsg = 1; }
            else {
                int start = 0;
                while (start < nchunk && nsg < kMaxSeg) {
                    const uint32_t limit = chunk_base(start) + stage_cap - (start == 0 ? seg_extra : 0u);
                    // inclusive prefix of chunk l / l + 32 beyond the limit?
                    const uint32_t in0 = ex0 + (lane < nchunk ? S.ctot[lane] : 0u), in1 = ex1 + (lane + 32 < nchunk ? S.ctot[lane + 32] : 0u);
                    const uint32_t f0 = __ballot_sync(0xffffffffu, lane < nchunk && in0 > limit);
                    const uint32_t f1 = __ballot_sync(0xffffffffu, lane + 32 < nchunk && in1 > limit);
                    unsigned long long F = ((unsigned long long)f1 << 32) | f0;
                    F &= ~((1ull << start) - 1ull);
                    int end = F ? __ffsll((long long)F) - 1 : nchunk;
                    if (end <= start) end = start + 1;             // cannot happen: one chunk always fits (checked at open)
                    sc[++nsg] = end;
                    start = end;
                }
                if (start < nchunk) { if (tid == 0) atomicOr(A.flags, 4u); sc[nsg] = nchunk; }
            }
            PHASE_MARK(1);

            // ---- class lists: class q's samples start at cb_q (classes back to back, warps in order inside a class, x order inside
            // a warp's share); every warp works the layout out for itself from the [class][warp] counts, then scatters the samples
            // it prepared. A sample's record offset is relative to its column segment.
            uint32_t klo, khi;
            {
                uint32_t before = 0, tot = 0;
                if (lane < NW) {
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(S.wcnt + lane * NW);
#pragma unroll
                    for (int w2 = 0; w2 < NW / 2; w2++) {
                        const uint32_t pr = row[w2], lo = pr & 0xFFFFu, hi = pr >> 16;
                        tot += lo + hi;
                        if (2 * w2 < warp) before += lo;
                        if (2 * w2 + 1 < warp) before += hi;
                    }
                }
                uint32_t inc = tot;
#pragma unroll
                for (int o = 1; o < NW; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                const uint32_t cbq = inc - tot, cstart = cbq + before;
                klo = __shfl_sync(0xffffffffu, cbq, warp & 31);
                khi = klo + __shfl_sync(0xffffffffu, tot, warp & 31);
#pragma unroll
                for (int k = 0; k < KCMAX; k++) {
                    const int c = warp * KC + k;
                    if (k < KC && c < nchunk) {
                        const uint32_t st0 = __shfl_sync(0xffffffffu, cstart, (scx[k] >> 16) & 31u);
                        int sg = 0;
                        while (sg + 1 < nsg && c >= sc[sg + 1]) sg++;
                        const uint32_t relc = (sg == 0 ? seg_extra : 0u) + chunk_base(c) - chunk_base(sc[sg]);
                        if (nsg > 1) {                   // samples of each class per column segment
                            const uint32_t m2 = __match_any_sync(0xffffffffu, scx[k] != 0xFFFFFFFFu ? scx[k] >> 16 : 32u + (uint32_t)lane);
                            if (scx[k] != 0xFFFFFFFFu && (m2 & lt) == 0) atomicAdd(&S.segcnt[sg * NW + (int)(scx[k] >> 16)], (uint32_t)__popc(m2));
                        }
                        if (scx[k] != 0xFFFFFFFFu) {
                            // what the symbol does to the 32 slots of its context (rangecoder::s, FFV1_RangeCoder.cpp:135-171): slot 0 zero
                            // flag | 1..10 exponent | 11..21 sign | 22..31 mantissa. Symbols with exponent > 9 use slots 10 and 31 more than
                            // once; they never take the mask-driven path.
                            const int v = sv[k];
                            const uint32_t a = (uint32_t)abs(v);
                            const int e = v ? 31 - __clz(a) : 0;
                            uint32_t um = 1u, bmk = v ? 0u : 1u;
                            if (v && e <= 9) {
                                const uint32_t le = (1u << e) - 1u;
                                um = 1u | ((((1u << (e + 1)) - 1u)) << 1) | (1u << (11 + e)) | (le << 22);
                                bmk = (le << 1) | ((v < 0 ? 1u : 0u) << (11 + e)) | ((a & le) << 22);
                            } else if (v) um = 0u;            // no mask for this one
                            const uint32_t o2 = 2u * (relc + sof[k]);
                            S.ent[st0 + spi[k]] = make_uint4(um, bmk, o2 | ((o2 + 4u * (uint32_t)e) << 16), ((uint32_t)v & 0x3FFFFu) | ((scx[k] & 0xFFFFu) << 18));
                        }
                    }
                }
            }
            __syncthreads();
            if (row_taken && tid == 0 && y + 2 < r1) issue_row(y + 2);      // the raw row has been consumed by everybody
            uint32_t kcur = klo;             // next sample of my class (the cursor runs on over the column segments)
            const uint32_t rc_base = ((uint32_t)(fs * A.band_rows + (y - r0)) * 3u + (uint32_t)(ps ? 1 + pl : 0)) * (uint32_t)A.nseg;
            for (int s = 0; s < nsg; s++) {
                const int c0 = sc[s], c1 = sc[s + 1];
                const uint32_t extra = s == 0 ? seg_extra : 0u;
                const uint32_t segbase = chunk_base(c0);
                const uint32_t seg_total = extra + chunk_base(c1) - segbase;
                // ---- S2 (K3): warp q = context class q. Its samples of this segment, in x order, 32 per batch.
                {
                    const uint32_t kend = nsg == 1 ? khi : kcur + S.segcnt[s * NW + warp];
                    // The samples of my class go one after the other, in list (= bitstream) order; lane s holds the state of slot s of
                    // the context being coded. What a symbol does to the 32 slots comes as the two masks S1 left in its entry (slot
                    // used / bin value), read from shared memory two samples ahead (same address for every lane); lanes whose slot
                    // the symbol does not use keep their state (predicated, not branched). A sample costs about twenty instructions
                    // whatever its context did before: consecutive samples of one context (low-noise and flat content: a handful of
                    // contexts take most of a row) pay one (bit, state) look-up of latency each, a change of context one store and
                    // one load of the 32 state bytes.
                    {
                        const int kA = lane == 0 ? 0 : isB ? lane : isD ? 2 : 23 - lane;     // record index of my slot's bin: kA + kE e, kE = 0 or 2
                        uint32_t abase = stage_a + 2u * (uint32_t)kA, hsel = (lane == 0 || isB) ? 0x4410u : 0x4432u, lbit = 1u << lane;
                        uint32_t tr0 = trans_a, tr1 = trans_a + 256u;
                        uint32_t srow = states_a + (uint32_t)lslot;
                        asm volatile("" : "+r"(abase), "+r"(hsel), "+r"(lbit), "+r"(tr0), "+r"(tr1), "+r"(srow));
                        // lanes without a slot (8-bit streams: 27 states per context) and the time before the first context use a
                        // landing byte of their own, so that the loads and stores of the state bytes need no further condition
                        const uint32_t spare = smem_addr(S.misc) + (uint32_t)warp * 64u + (uint32_t)lane;
                        uint32_t cur = 0xFFFFFFFFu, st = 128u, sp = spare;
                        auto flush = [&]() { sts_u8(sp, st); cur = 0xFFFFFFFFu; sp = spare; };
                        // kWide: the batch holds symbols with exponent > 9 (no masks: entry.x == 0), coded out of line
                        auto sample_step = [&](const uint4& E, auto wide_tag) {
                            constexpr bool kWide = decltype(wide_tag)::value;
                            const uint32_t cxs = E.w >> 18;
                            const uint32_t spn = lane_has_slot ? srow + cxs * (uint32_t)kRow : spare;
                            // change of context (warp-wide, predicated): the state bytes go back, the new context's come in
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, %4;\n\t@p st.shared.u8 [%1], %0;\n\t@p ld.shared.u8 %0, [%2];\n\t}"
                                         : "+r"(st) : "r"(sp), "r"(spn), "r"(cxs), "r"(cur) : "memory");
                            sp = spn;
                            cur = cxs;
                            if (!kWide || E.x != 0u) {
                                const uint32_t used = E.x & lbit, bitm = E.y & lbit;
                                const uint32_t ta = (bitm ? tr1 : tr0) + st;
                                const uint32_t rec = bitm ? 255u + st : 255u - st;
                                const uint32_t dst = abase + __byte_perm(E.z, 0, hsel);
                                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.u8 %0, [%4];\n\t@p st.shared.u16 [%1], %2;\n\t}"
                                             : "+r"(st) : "r"(dst), "h"((uint16_t)rec), "r"(used), "r"(ta) : "memory");
                            } else {
                                st = wide_symbol(E, st, lane, stage_a, trans_a);
                            }
                        };
                        while (kcur < kend) {
                            const uint32_t nbt = min(32u, kend - kcur);
                            const bool have = (uint32_t)lane < nbt;
                            uint4 en = make_uint4(1u, 1u, 0u, 0u);
                            if (have) en = S.ent[kcur + (uint32_t)lane];
                            const uint32_t cx0 = __shfl_sync(0xffffffffu, en.w >> 18, 0);
                            // runs of zeros: when every sample of the batch has residual 0 and they share a context (flat areas, mattes,
                            // black frames), the k-th of them sees state one_state^k(st) of slot 0 and nothing else moves: all of them
                            // at once, through the tables of one_state^(2^i)
                            if (__all_sync(0xffffffffu, !have || ((en.w & 0x3FFFFu) == 0u && (en.w >> 18) == cx0))) {
                                if (cur == cx0) flush();                               // the zero run works on the state row in shared memory
                                __syncwarp();
                                const uint32_t sa = states_a + cx0 * (uint32_t)kRow;          // slot 0 = byte 0 of the row
                                uint32_t z0 = lds_u8_volatile(sa);
                                __syncwarp();
#pragma unroll
                                for (int kk = 0; kk < 5; kk++)
                                    if ((lane >> kk) & 1) z0 = lds_u8(tpow_a + kk * 256 + z0);
                                if (have) sts_u16(stage_a + (en.z & 0xFFFFu), z0 + 255u);                             // q = st - 1, bit 1
                                if ((uint32_t)lane == nbt - 1u) sts_u8(sa, lds_u8(tpow_a + z0));
                                __syncwarp();
                            } else {
                                // each entry is asked for one sample ahead of its use (right after its register set has been consumed)
                                const uint32_t ea = smem_addr(S.ent) + kcur * 16u;
                                const bool anywide = __any_sync(0xffffffffu, have && en.x == 0u);
                                auto walk = [&](auto wide_tag) {
                                    uint4 E0 = lds_v4(ea), E1 = lds_v4(ea + 16u);          // (past the end of the list: still shared memory, unused)
                                    uint32_t off = 32u;
                                    const uint32_t pairs = nbt >> 1;
                                    for (uint32_t jj = 0; jj < pairs; jj++, off += 32u) {
                                        sample_step(E0, wide_tag);
                                        E0 = lds_v4(ea + off);
                                        sample_step(E1, wide_tag);
                                        E1 = lds_v4(ea + off + 16u);
                                    }
                                    if (nbt & 1u) sample_step(E0, wide_tag);
                                };
                                if (anywide) walk(std::true_type{}); else walk(std::false_type{});
                            }
                            kcur += nbt;
                        }
                        flush();
                        __syncwarp();
                    }
                }
                // no-op records up to the next whole block
                const uint32_t padded = (seg_total + (uint32_t)kBlockRecs - 1u) & ~((uint32_t)kBlockRecs - 1u);
                if ((uint32_t)tid < padded - seg_total) S.stage[seg_total + tid] = (uint16_t)kNopRec;
                __syncthreads();
                PHASE_MARK(2);
                // ---- S3: staged records -> global: q bytes and bit-plane, whole blocks, coalesced
                {
                    const uint32_t ngrp = padded >> 4;          // groups of 16 records; a multiple of 8
                    uint4* dq = reinterpret_cast<uint4*>(gq + pos);
                    uint32_t* db = reinterpret_cast<uint32_t*>(gb + (pos >> 3));
                    const uint4* sv = reinterpret_cast<const uint4*>(S.stage);
                    for (uint32_t g0 = (uint32_t)warp * 32u; g0 < ngrp; g0 += kModelThreads) {
                        const uint32_t gi = g0 + lane;
                        const bool ok = gi < ngrp;
                        uint32_t bits16 = 0;
                        if (ok) {
                            const uint4 a = sv[2 * gi], b = sv[2 * gi + 1];
                            uint4 qv;
                            qv.x = __byte_perm(a.x, a.y, 0x6420); qv.y = __byte_perm(a.z, a.w, 0x6420);
                            qv.z = __byte_perm(b.x, b.y, 0x6420); qv.w = __byte_perm(b.z, b.w, 0x6420);
                            dq[gi] = qv;
                            const uint32_t h0 = __byte_perm(a.x, a.y, 0x7531), h1 = __byte_perm(a.z, a.w, 0x7531);
                            const uint32_t h2 = __byte_perm(b.x, b.y, 0x7531), h3 = __byte_perm(b.z, b.w, 0x7531);
                            bits16 = ((h0 * 0x01020408u) >> 24) | (((h1 * 0x01020408u) >> 24) << 4) |
                                     (((h2 * 0x01020408u) >> 24) << 8) | (((h3 * 0x01020408u) >> 24) << 12);
                        }
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, bits16, 1);
                        if (ok && !(lane & 1)) db[gi >> 1] = bits16 | (other << 16);
                    }
                    if (tid == 0) A.rowcnt[rc_base + s] = seg_total;
                    pos += padded;
                    bins_total += seg_total - extra;
                }
                if (s + 1 < nsg) __syncthreads();
                PHASE_MARK(3);
            }
            if (tid == 0) for (int s = nsg; s < A.nseg; s++) A.rowcnt[rc_base + s] = 0;
            seg_extra = 0;
        }
    }
#ifdef B200_PHASE_TIMING
    if (tid == 0) for (int k = 0; k < 8; k++) atomicAdd(reinterpret_cast<unsigned long long*>(A.flags + 16) + k, (unsigned long long)ph[k]);
#endif
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(A.flags + 2), bins_total);
    __syncthreads();
    if (r1 < g.h) {   // carry the states to the next band: one bulk copy out of shared memory (its completion is awaited by
                      // k_model_loop before the table is touched again)
        fence_async_smem();
        __syncthreads();
        if (tid == 0) bulk_s2g(save, states_a, (uint32_t)state_bytes);
    }
}

// Persistent grid: every CTA pulls (frame, slice, plane-set) items from a counter until none is left, the Cb/Cr items (twice
// the work) first. The grid is the number of SMs minus the ones left to k_range and k_emit, which must never wait behind
// queued model CTAs (launch_model).
template <bool kCompact>
__device__ __forceinline__ void k_model_loop(const EncArgs& A, int band, int nframes, uint8_t* smem_raw) {
    __shared__ int s_item;
    const int nfs = nframes * A.nslices, nitems = nfs * 2;
    const int tid = threadIdx.x;
    {   // what every item needs, once per CTA: the quantisation tables, the zero-run and (bit, state) tables, the mbarriers
        const ModelSmem S = carve(smem_raw, A.nctx, kCompact ? 28 : 32, A.wmax, 2);
        for (int i = tid; i < 5 * 256; i += kModelThreads) S.qtab[i] = A.qtab[i];
        for (int i = tid; i < 5 * 256; i += kModelThreads) S.tpow[i] = A.tpow[i];
        // next state by (bit, state): the only dependency from sample to sample of a context is one look-up in this table
        for (int i = tid; i < 512; i += kModelThreads) {
            const int st = i & 255;
            S.trans[i] = st == 0 ? (uint8_t)0 : (i >> 8) ? A.t1q[st - 1] : (uint8_t)(256 - A.t1q[255 - st]);
        }
        if (tid == 0) {
            mbar_init(smem_addr(S.mbar), 1);
            mbar_init(smem_addr(S.mbar) + 8u, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    uint32_t ph_state = 0, ph_row = 0;
    for (;;) {
        if (tid == 0) bulk_wait_read();        // the state table of the previous item has left shared memory
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(A.work_ctr + band, 1u);
        __syncthreads();
        const int wk = s_item;
        if (wk >= nitems) break;
        const int ps = wk < nfs ? 1 : 0, fsl = wk < nfs ? wk : wk - nfs;
        k_model_body<kCompact>(A, band, nframes, smem_raw, fsl % A.nslices, ps, fsl / A.nslices, ph_state, ph_row);
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the last table saved has reached global memory
}
__global__ void __launch_bounds__(kModelThreads, 1) k_model(const __grid_constant__ EncArgs A, int band, int nframes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    k_model_loop<false>(A, band, nframes, smem_raw);
}
__global__ void __launch_bounds__(kModelThreads, 1) k_model_compact(const __grid_constant__ EncArgs A, int band, int nframes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    k_model_loop<true>(A, band, nframes, smem_raw);
}

// ------------------------------------------------------------------------------------------------------------------
// The range coder, split in two.
//
// A binary range coder's output is one big number: every bin with value 1 adds (range - range1) at the byte position the
// coder has reached (FFV1_RangeCoder.cpp:71-102 seen from the encoder side: low += range - range1; on renormalisation the
// top byte of the 16-bit window leaves). Only `range` and the number of renormalisations form a serial recurrence.
//
//   k_range  one lane per (frame, slice): runs just that recurrence — range' = renorm((range * sp + c) >> 8) and the
//            shift count — over the slice's records, branch-free, records prefetched four 16-record groups ahead into
//            registers (no shared memory: its CTAs run beside k_model's). Every 64 records it leaves a checkpoint
//            (range, bytes so far).
//   k_emit   one thread per 64 records, fully parallel: replays them from the checkpoint with the complete coder step and
//            adds the contribution into the slice's byte stream, held as big-endian 32-bit words, with atomic adds
//            (neighbouring pieces overlap by the 16-bit window and by carries; addition commutes).
//
// Slice header bins are ordinary records at the head of the Y stream (written by k_model), the terminator bin only moves
// `range` (k_range), and the coder's final flush is a single +0xFF (k_range). CRC and footer: k_pack.

// add v into big-endian word idx of the slice stream, propagating wrap-arounds towards the front
__device__ __forceinline__ void stream_add(uint32_t* W, int64_t idx, uint32_t v) {
    while (v && idx >= 0) {
        const uint32_t old = atomicAdd(W + idx, v);
        v = (old + v < old) ? 1u : 0u;
        idx--;
    }
}

__device__ __forceinline__ void range_step(uint32_t sp, uint32_t c, uint32_t& range, uint32_t& cnt) {
    const uint32_t x = range * sp + c;     // bit 0: range - ((range*state)>>8) == (range*(256-state)+255)>>8, so c = 255
    const bool sh = x < 0x10000u;
    range = sh ? (x & 0xFFFF00u) : (x >> 8);
    cnt += sh;
}

constexpr int kCkptRecs = 64;

// 0xFF when the most significant bit of byte `byte` of v is set, else 0 (prmt's sign-replication mode)
template <int kByte>
__device__ __forceinline__ uint32_t msb_mask(uint32_t v) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(0u), "r"(0x4440u | 8u | (uint32_t)kByte));
    return d;
}
template <int kByte>
__device__ __forceinline__ uint32_t get_byte(uint32_t v) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(0u), "r"(0x4440u | (uint32_t)kByte));
    return d;
}
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_stream_u16(const void* p) {
    uint16_t v;
    asm volatile("ld.global.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}

// ---- k_range: the records reach the coder lanes through shared memory, asynchronously.
// Every plane-row segment is padded to whole 128-record blocks, so the 32 lanes of a warp (32 slices) cross block boundaries
// in the same iteration although their streams differ. At a boundary every lane asks for the block it will need two blocks
// later with nine 16-byte cp.async copies (128 q bytes + 16 bit-plane bytes, global -> shared without passing through
// registers), and waits for the one it is about to read, which was asked for 256 records earlier. The recurrence itself then
// only reads shared memory: one 16-byte and one 2-byte load per 16 records, no global-memory latency in the chain, and the
// cursor over the row segments runs once per 128 records instead of once per 16.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16v(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}

constexpr int kRangeWarps = 4;                       // one per scheduler
#ifndef B200_RANGE_SLOTS
#define B200_RANGE_SLOTS 4
#endif
constexpr int kRangeSlots = B200_RANGE_SLOTS;        // blocks in a lane's ring: one being read, kRangeSlots - 2 on their way, one just read
constexpr int kRangeAhead = kRangeSlots - 2;
#ifndef B200_RANGE_UNROLL
#define B200_RANGE_UNROLL 1
#endif
constexpr int kRangeUnroll = B200_RANGE_UNROLL;      // 16-record groups of a block unrolled together (the whole kernel unrolled is 100 KB of code)
constexpr uint32_t kRangeSlotStride = 144 + 16;      // 128 q bytes + 16 bit-plane bytes (+ pad: spreads the lanes over the banks)
constexpr uint32_t kRangeWarpBytes = kRangeSlots * 32 * kRangeSlotStride;
// the rings take 80 KB; the CTA asks for more than half an SM's shared memory so that it has the SM to itself: a warp that
// shares its scheduler with others of its kind slows down by the round-robin factor
constexpr size_t kRangeSmem = kRangeSlots > 4 ? 200 * 1024 : 120 * 1024;
static_assert(kRangeWarps * kRangeWarpBytes <= kRangeSmem, "k_range rings");

__global__ void __launch_bounds__(32 * kRangeWarps) k_range(const __grid_constant__ EncArgs A, int band, int nframes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = (blockIdx.x * kRangeWarps + warp) * 32 + lane;
    const bool in_range = gid < nframes * A.nslices;
    const int slice = in_range ? gid % A.nslices : 0;
    const SliceGeom g = A.geom[slice];
    const int r0 = band * A.band_rows;
    const bool valid = in_range && r0 < g.h;
    const int r1 = min(r0 + A.band_rows, g.h);
    const int nseg = A.nseg;
    const int nt = valid ? (r1 - r0) * 3 * nseg : 0;           // segments in bitstream order: row, plane, column segment
    const size_t gsafe = in_range ? (size_t)gid : 0;
    const uint32_t* rc = A.rowcnt + gsafe * A.band_rows * 3 * nseg;

    // 128-record blocks this lane consumes in this band, per stream (Y: plane 0, C: planes 1 and 2)
    uint32_t totY = 0, totC = 0;
    for (int t = 0; t < nt; t++) {
        const uint32_t nbk = (rc[t] + 127u) >> 7;
        if (((t / nseg) % 3) != 0) totC += nbk; else totY += nbk;
    }
    uint32_t maxb = totY + totC;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxb = max(maxb, __shfl_xor_sync(0xffffffffu, maxb, o));

    uint32_t range = 0xFF00u, cnt = 0;
    if (valid && band > 0) { const CoderState s = A.cstate[gid]; range = s.range; cnt = s.pos; }

    const uint8_t* gqY = A.qY + gsafe * A.capY;
    const uint8_t* gqC = A.qC + gsafe * A.capC;
    const uint8_t* gbY = A.bY + gsafe * (A.capY >> 3);
    const uint8_t* gbC = A.bC + gsafe * (A.capC >> 3);
    uint2* kY = A.ckptY + gsafe * (A.capY / kCkptRecs);
    uint2* kC = A.ckptC + gsafe * (A.capC / kCkptRecs);

    // this lane's ring: slot k at ring + k * 32 * kRangeSlotStride
    const uint32_t ring = smem_addr(smem_raw) + (uint32_t)warp * kRangeWarpBytes + (uint32_t)lane * kRangeSlotStride;
    auto slot_at = [&](uint32_t k) { return ring + k * 32u * kRangeSlotStride; };

    // block cursor (fetch side): the lane's blocks in bitstream order
    int t = -1;
    uint32_t rem = 0, bkY = 0, bkC = 0, isC = 0;
    uint32_t cnt_pref = nt > 0 ? rc[0] : 0u;
    // asks for the lane's next block into slot k; returns its tag: bit 31 = a block, bit 30 = C stream, low bits = block index
    auto fetch_block = [&](uint32_t k) -> uint32_t {
        while (rem == 0 && t + 1 < nt) {
            t++;
            const uint32_t c = cnt_pref;
            cnt_pref = t + 1 < nt ? rc[t + 1] : 0u;
            rem = (c + 127u) >> 7;
            isC = ((t / nseg) % 3) != 0;
        }
        uint32_t tag = 0;
        if (rem) {
            rem--;
            const uint32_t bi = isC ? bkC : bkY;
            const uint8_t* q = (isC ? gqC : gqY) + (size_t)bi * 128u;
            const uint8_t* b = (isC ? gbC : gbY) + (size_t)bi * 16u;
            const uint32_t dst = slot_at(k);
#pragma unroll
            for (int i = 0; i < 8; i++) cp_async16(dst + i * 16, q + i * 16);
            cp_async16(dst + 128, b);
            if (isC) bkC = bi + 1; else bkY = bi + 1;
            tag = 0x80000000u | (isC << 30) | bi;
        }
        cp_async_commit();                      // an empty group when the lane has run out of blocks: the counts stay in step
        return tag;
    };
    uint32_t tags[kRangeSlots];                 // tag of the block in each slot (compile-time indexed: the block loop is unrolled by 4)
#pragma unroll
    for (int i = 0; i < kRangeSlots; i++) tags[i] = 0;
#pragma unroll
    for (int i = 0; i < kRangeAhead; i++) tags[i] = fetch_block((uint32_t)i);

    for (uint32_t b0 = 0; b0 < maxb; b0 += kRangeSlots) {
#pragma unroll
        for (int k = 0; k < kRangeSlots; k++) {
            // block b0 + k lives in slot k; ask for block b0 + k + 2 (its slot held block b0 + k - 2: read long ago), then wait
            // until at most the two newest requests are pending
            tags[(k + kRangeAhead) % kRangeSlots] = fetch_block((uint32_t)((k + kRangeAhead) % kRangeSlots));
            cp_async_wait<kRangeAhead>();
            const uint32_t tag = tags[k];
            const bool have = (tag & 0x80000000u) != 0;
            const uint32_t sl = slot_at((uint32_t)k);
            uint2* ck = ((tag >> 30) & 1u ? kC : kY) + (size_t)(tag & 0x3FFFFFFFu) * 2u;
#pragma unroll(kRangeUnroll)
            for (int gI = 0; gI < 8; gI++) {
                uint4 q = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);      // no-op records for a lane without a block
                uint32_t nbits = 0xFFFFFFFFu;
                if (have) {
                    q = lds_v4(sl + gI * 16);
                    nbits = ~lds_u16v(sl + 128 + gI * 2);
                    if ((gI & 3) == 0) ck[gI >> 2] = make_uint2(range, cnt);                   // checkpoint every 64 records
                }
                const uint32_t ww[4] = {q.x, q.y, q.z, q.w};
#define RSTEP(k) range_step(get_byte<((k) & 3)>(ww[(k) >> 2]) + 1u, msb_mask<((k) >> 3)>(nbits << (7 - ((k) & 7))), range, cnt);   /* c = 255 when the bit is 0 */
                RSTEP(0) RSTEP(1) RSTEP(2) RSTEP(3) RSTEP(4) RSTEP(5) RSTEP(6) RSTEP(7)
                RSTEP(8) RSTEP(9) RSTEP(10) RSTEP(11) RSTEP(12) RSTEP(13) RSTEP(14) RSTEP(15)
#undef RSTEP
            }
        }
    }
    cp_async_wait<0>();
    if (in_range) { A.used[gsafe * 2] = totY << 1; A.used[gsafe * 2 + 1] = totC << 1; }
    if (valid) {
        if (r1 == g.h) {
            range_step(127u, 255u, range, cnt);                    // terminator: state 129, bit 0 (FFV1_Slice.cpp:334-343)
            // final flush of the coder (so that the decoder's BytesUsed() lands on the end, FFV1_Slice.cpp:297-299):
            // low += 0xFF in the current 16-bit window (stream bytes cnt, cnt+1), then two forced renormalisations
            uint32_t* W = reinterpret_cast<uint32_t*>(A.scratch + gsafe * A.slice_cap);
            const uint32_t q = cnt + 1;
            if ((size_t)q + 8 < A.slice_cap) stream_add(W, q >> 2, 0xFFu << ((3 - (q & 3)) * 8));
            else atomicOr(A.flags, 1u);
            A.slice_size[gid] = cnt + 1;                           // cnt + 2 bytes produced; the last stays inside the coder
        } else {
            CoderState s;
            s.range = range; s.pos = cnt;
            A.cstate[gid] = s;
        }
    }
}

// k_emit: grid (frame*slice, stream Y/C, pieces of 256 x 64 records): the large dimension in x, which has no 65535 limit
constexpr int kEmitThreads = 256;
constexpr int kEmitWords = 22;          // local window: word 0 spare, then stream words (c0>>2)-1 ... (c0>>2)+19

__global__ void __launch_bounds__(kEmitThreads) k_emit(const __grid_constant__ EncArgs A) {
    __shared__ uint32_t s_loc[kEmitWords][kEmitThreads];
    const int gid = blockIdx.x, stream = blockIdx.y, tid = threadIdx.x;
    const uint32_t used = A.used[(size_t)gid * 2 + stream];
    if (blockIdx.z * kEmitThreads >= used) return;
    const uint32_t blk = blockIdx.z * kEmitThreads + tid;
    if (blk >= used) return;
    const size_t cap = stream ? A.capC : A.capY;
    const uint4* src = reinterpret_cast<const uint4*>((stream ? A.qC : A.qY) + (size_t)gid * cap) + (size_t)blk * 4;
    const uint2 bw = *(reinterpret_cast<const uint2*>((stream ? A.bC : A.bY) + (size_t)gid * (cap >> 3)) + blk);
    const uint2 ck = (stream ? A.ckptC : A.ckptY)[(size_t)gid * (cap / kCkptRecs) + blk];
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; u++) q[u] = __ldg(src + u);
#pragma unroll
    for (int i = 0; i < kEmitWords; i++) s_loc[i][tid] = 0;

    uint32_t range = ck.x;
    const uint32_t c0 = ck.y;
    // byte `lb` of the local window is stream byte 4*((c0>>2)-1) + lb - 4; the piece starts at local byte 8 + (c0&3)
    uint32_t lpos = 8 + (c0 & 3);           // local byte index of the coder window's high byte
    uint64_t acc = 0;                       // bits 0..15 window, then pending bytes, then carry
    uint32_t nsh = 0;
    auto flush = [&](uint32_t extra) {
        // the value above the window (pending bytes + carry) has its LSB at the end of local byte lpos + nsh/8 - 1;
        // `extra` = 16 also retires the window itself (end of the piece)
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
    auto step = [&](uint32_t qb, uint32_t bit) {
        const uint32_t sp = qb + 1u;
        const uint32_t x = range * sp + (bit ? 0u : 255u);
        const uint32_t nr = x >> 8;
        acc += bit ? (uint64_t)(range - nr) : 0ull;
        const uint32_t sh = x < 0x10000u ? 8u : 0u;
        range = nr << sh;
        acc <<= sh;
        nsh += sh;
    };
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t bits = (u < 2 ? bw.x : bw.y) >> ((u & 1) * 16);
        const uint32_t ww[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            step((ww[k >> 2] >> ((k & 3) * 8)) & 255u, (bits >> k) & 1u);
            if ((k & 3) == 3) flush(0);
        }
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
    // zero on consume: k_emit accumulates into the slice stream with atomic adds, so the next encode call needs it zeroed.
    // What this call touched ends a few words beyond byte n (k_emit's local window, k_range's flush): clearing it here costs
    // a write of the bytes just read instead of a memset of the worst-case capacity before every call.
    {
        uint4* z = reinterpret_cast<uint4*>(A.scratch + (size_t)gid * A.slice_cap);
        const uint32_t nz = (uint32_t)(min((size_t)n + 160, A.slice_cap) + 15) >> 4;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t i = tid; i < nz; i += kPackThreads) z[i] = zero;
    }
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
typedef void (*model_fn)(const EncArgs, int, int);
static model_fn pick_model(const EncArgs& a) { return a.sstride != 32 ? k_model_compact : k_model; }

cudaError_t configure_kernels(const EncArgs& a) {
    const size_t need = model_smem_bytes(a.nctx, a.sstride, a.wmax, 2, a.stage_cap);
    // every kernel asks for the same L1 / shared-memory split as k_model: an SM cannot hold CTAs of two kernels that want
    // different splits, and k_range / k_emit / k_pack must run beside k_model's CTAs, not after them
    cudaFuncSetAttribute(k_range, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_range, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRangeSmem);
    cudaFuncSetAttribute(k_emit, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_pack, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pick_model(a), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaFuncSetAttribute(pick_model(a), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
}

cudaError_t launch_model(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    const size_t smem = model_smem_bytes(a.nctx, a.sstride, a.wmax, 2, a.stage_cap);
    int grid = a.model_ctas;
    if (grid > nframes * a.nslices * 2) grid = nframes * a.nslices * 2;
    pick_model(a)<<<grid, kModelThreads, smem, s>>>(a, band, nframes);
    return cudaGetLastError();
}

cudaError_t launch_range(const EncArgs& a, int band, int nframes, cudaStream_t s) {
    const int n = nframes * a.nslices;
    const int warps = (n + 31) / 32;
    // one lane per (frame, slice); CTAs of four warps, one per scheduler. The record rings (168 KB) keep a CTA alone on its
    // SM: a warp that shares its scheduler with others of its kind slows down by the round-robin factor.
    // dynamic shared memory: the rings (80 KB), or more when the caller wants fewer CTAs per SM (a.range_smem)
    size_t smem = kRangeWarps * kRangeWarpBytes;
    if ((size_t)a.range_smem > smem) smem = (size_t)a.range_smem;
    if (smem > kRangeSmem) smem = kRangeSmem;
    k_range<<<(warps + kRangeWarps - 1) / kRangeWarps, 32 * kRangeWarps, smem, s>>>(a, band, nframes);
    return cudaGetLastError();
}

cudaError_t launch_emit(const EncArgs& a, int nframes, cudaStream_t s) {
    int n = nframes * a.nslices;
    const unsigned nblk = (unsigned)((a.capC / kCkptRecs) + kEmitThreads - 1) / kEmitThreads;
    k_emit<<<dim3(n, 2, nblk), kEmitThreads, 0, s>>>(a);
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
