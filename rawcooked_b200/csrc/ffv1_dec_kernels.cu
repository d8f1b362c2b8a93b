// sm_100a kernels of the B200 FFV1 decoder (`--check` side; interface in ffv1_dec.h, C ABI in include/b200dec.h).
//
// The decoder is serial per slice in both the context model and the range coder: every decoded sample is the left
// neighbour of the next one's predictor and context. The parallelism is therefore slices x frames in flight: every slice
// of the batch is decoded by ONE lane, `spw` slices share a warp (their lanes step through their streams sample by sample in
// lockstep, so the warp issues one instruction stream for spw slices), and a finished row of any of them is turned into
// file bytes by all 32 lanes (inverse RCT, byte layout, store and/or compare with the source payload).
// Measured (B200, 4K 16-bit grain, 24 slices): a lane needs ~3 200 cycles per sample (~680 dependent instructions, one
// every ~4.6 cycles: the chain of (low, range) through ~17 bins, plus one random 32-byte read of the context's states that no
// cache holds, because the quantised context of noisy 16-bit material is spread over all 10 126 rows); throughput therefore
// comes from warps in flight: 58 fps at 128 frames in flight (2 slices per warp), 104 fps at 256 (4 per warp). Capping the
// registers for more resident warps spills the coder state and loses (80 registers: 36 fps), hence B200_DEC_MIN_CTAS = 4.
//
//   packet -> slices (tail walk)        /root/reference/Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:166-197        k_dec_index
//   slice CRC, keyframe bin, header     FFV1_Slice.cpp:210-260, :113-177                                     k_decode
//   rows: LineThenPlane / Line          FFV1_Slice.cpp:406-472 (borders :430-433), predict / context :21-93
//   symbol, bin                         FFV1_RangeCoder.cpp:71-102 (b), :135-171 (s), :105-132 (u), :51-62 (BytesUsed, IsUnderrun)
//   terminator, junk, footer            FFV1_Slice.cpp:335-346, :294-313
//   inverse RCT + byte layouts          Source/Lib/Transform/Transform.cpp:29-37, :70-420
#include "ffv1_dec.h"

#include "../../include/b200dec.h"

namespace b200 {
namespace {

__device__ __forceinline__ uint32_t bswap32d(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint32_t bswap16d(uint32_t v) { return __byte_perm(v, 0, 0x4401); }
__device__ __forceinline__ uint32_t smem_a(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int median3d(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

#ifndef B200_DEC_MIN_CTAS
#define B200_DEC_MIN_CTAS 4
#endif
constexpr int kRowStride = 48;    // bytes between the staged state rows of two lanes (32 used): banks 12 apart, no conflict up to 8 lanes

// ------------------------------------------------------------------------------------------------------------------
// k_dec_index: one thread per packet walks the slice tails from the end of the packet (FFV1_Frame.cpp:166-197)
__global__ void k_dec_index(const __grid_constant__ DecArgs A, int nframes) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint8_t* p = A.packets + A.pkt_off[f];
    uint64_t pos = A.pkt_len[f];
    const uint32_t tail = (uint32_t)A.tail;
    uint32_t st = 0;
    int k = 0;
    while (pos) {
        if (pos < tail || k >= A.nslices) { st |= B200_DEC_BAD_TAIL; break; }
        const uint8_t* t = p + pos - tail;
        uint64_t size = ((uint32_t)t[0] << 16) | ((uint32_t)t[1] << 8) | t[2];
        size += tail;
        if (size > pos) { st |= B200_DEC_BAD_TAIL; break; }
        pos -= size;
        A.sl_off[(size_t)f * A.nslices + k] = A.pkt_off[f] + pos;
        A.sl_size[(size_t)f * A.nslices + k] = (uint32_t)size;
        k++;
    }
    if (k != A.nslices) st |= B200_DEC_BAD_TAIL;
    for (; k < A.nslices; k++) A.sl_size[(size_t)f * A.nslices + k] = 0;
    A.status[f] = st;
    A.mismatch[f] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// range decoder of one lane (rangecoder, FFV1_RangeCoder.cpp:25-102): `low` = Current, `range` = Mask; the byte a
// renormalisation needs is consumed when the next bin is asked for, as the reference does, so that the byte accounting at the
// end of the slice (BytesUsed) is the reference's. The byte stream is read as aligned 32-bit words into a 64-bit window
// (hi:lo, next byte on top, `nav` unread bytes), one word (`nxt`) ahead of the window, so that no load sits in the coder's
// dependency chain and a renormalisation is five predicated instructions without a branch: the window is topped up to at
// least four bytes before every group of at most four bins (a bin consumes at most one byte). Words are byte-swapped and the
// bytes at or beyond the end of the range-coded data zeroed when they are loaded ("last byte assumed to be 0x00",
// FFV1_RangeCoder.cpp:79-84). A stream that runs dry keeps reading zeros and is flagged at the end of the slice (the
// reference answers 0 to every further bin instead; both report the slice as broken).
struct Rc {
    const uint32_t* wp;     // address of the word in `nxt`
    uint32_t hi, lo, nxt, nav;
    int bnext;              // stream offset of the first byte of `nxt` (the first word may start before the stream)
    uint32_t low, range, end;
};
__device__ __forceinline__ uint32_t rc_word(const uint32_t* p, int off, uint32_t end) {     // word at stream offset off
    const int rem = (int)end - off;
    uint32_t w = 0;
    if (rem > 0) {
        w = bswap32d(*p);
        if (rem < 4) w &= 0xFFFFFFFFu << (8 * (4 - rem));
    }
    return w;
}
// nav <= 4: append `nxt` below the unread bytes, fetch the word after it
__device__ __forceinline__ void rc_refill(Rc& c) {
    const unsigned long long add = (unsigned long long)c.nxt << (32u - 8u * c.nav);
    c.hi |= (uint32_t)(add >> 32);
    c.lo |= (uint32_t)add;
    c.nav += 4;
    c.wp++;
    c.bnext += 4;
    c.nxt = rc_word(c.wp, c.bnext, c.end);
}
__device__ __forceinline__ void rc_ensure4(Rc& c) { if (c.nav < 4u) rc_refill(c); }
__device__ __forceinline__ uint32_t rc_consumed(const Rc& c) { return (uint32_t)(c.bnext - (int)c.nav); }   // Buffer_Cur - Buffer_Beg
__device__ __forceinline__ void rc_take_byte(Rc& c, bool on) {     // low = low << 8 | next byte
    if (on) {
        c.low = __funnelshift_l(c.hi, c.low, 8);
        c.hi = __funnelshift_l(c.lo, c.hi, 8);
        c.lo <<= 8;
        c.nav--;
    }
}
__device__ __forceinline__ void rc_init(Rc& c, const uint8_t* buf, uint32_t size) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(buf);
    c.wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const int skip = (int)(a & 3);
    c.end = size;
    c.hi = rc_word(c.wp, -skip, size) << (8 * skip);
    c.lo = 0;
    c.nav = 4 - skip;
    c.wp++;
    c.bnext = 4 - skip;
    c.nxt = rc_word(c.wp, c.bnext, size);
    rc_refill(c);                        // nav = 5..8
    c.low = 0;
    rc_take_byte(c, true);               // AssignBuffer: Current = first byte, Mask = 0xFF, Buffer_Cur = 1
    c.range = 0xFF;
}
__device__ __forceinline__ void rc_dead(Rc& c) { c.wp = nullptr; c.hi = c.lo = c.nxt = 0; c.nav = 8; c.bnext = 8; c.low = 0; c.range = 0xFF00; c.end = 0; }
// the renormalisation at the head of rangecoder::b (FFV1_RangeCoder.cpp:73-88), under a predicate; needs nav >= 1
__device__ __forceinline__ void rc_renorm(Rc& c, bool on) {
    const bool rn = on && c.range < 0x100u;
    rc_take_byte(c, rn);
    if (rn) c.range <<= 8;
}
// the decision; tt = (next state after 0) | (next state after 1) << 8 of the state st. Returns the bit, ns = the next state
__device__ __forceinline__ uint32_t rc_decide(Rc& c, uint32_t st, uint32_t tt, bool on, uint32_t& ns) {
    const uint32_t r1 = (c.range * st) >> 8;
    const uint32_t r0 = c.range - r1;
    const bool bit = c.low >= r0;
    if (on) {
        c.low = bit ? c.low - r0 : c.low;
        c.range = bit ? r1 : r0;
    }
    ns = bit ? tt >> 8 : tt & 0xFFu;
    return bit ? 1u : 0u;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
// one bin on the adaptive state at shared-memory address sa (slice header, terminator, symbols with exponent >= 10)
__device__ __forceinline__ uint32_t rc_bin(Rc& c, uint32_t sa, uint32_t tt_a) {
    const uint32_t st = lds8(sa);
    rc_ensure4(c);
    rc_renorm(c, true);
    uint32_t ns;
    const uint32_t bit = rc_decide(c, st, lds16(tt_a + 2u * st), true, ns);
    sts8(sa, ns);
    return bit;
}
// rangecoder::u (FFV1_RangeCoder.cpp:105-132) on the 32 states at row_a (slice header)
__device__ __forceinline__ uint32_t rc_u(Rc& c, uint32_t row_a, uint32_t tt_a, bool& broken) {
    if (rc_bin(c, row_a, tt_a)) return 0;
    int e = 0;
    while (rc_bin(c, row_a + 1 + min(e, 9), tt_a)) {
        if (++e > 31) { broken = true; return 0; }
    }
    uint32_t a = 1;
    for (int i = e - 1; i >= 0; i--) a = (a << 1) | rc_bin(c, row_a + 22 + min(i, 9), tt_a);
    return a;
}

// The 32 states of a context in registers. A symbol with exponent <= 9 touches every slot at most once
// (rangecoder::s, FFV1_RangeCoder.cpp:135-171: slot 0, exponent slots 1 + i, mantissa slots 22 + i, sign slot 11 + e), so the
// states are read from `in` and the updated ones collected in `out`, and the transition pairs of the slots a symbol may use
// are looked up before its first bin: neither a state read nor a table look-up is left in the dependency chain of
// (low, range), and no instruction waits on shared memory. The slots are compile-time constants. All 32 lanes run the same
// instruction stream: a lane takes part in a step under a predicate, lanes leave the exponent steps one by one and wait for
// the others where the paths join, the mantissa steps start at the largest exponent among the lanes. Lanes without a slice
// compute on zeros and touch no memory.
struct Row { uint32_t in[8], out[8]; };
template <int SLOT>
__device__ __forceinline__ uint32_t row_get(const Row& R) { return (R.in[SLOT >> 2] >> (8 * (SLOT & 3))) & 0xFFu; }
// byte (SLOT & 3) of out[SLOT >> 2] = v when on
template <int SLOT>
__device__ __forceinline__ void row_put(Row& R, uint32_t v, bool on) {
    constexpr uint32_t ins = (SLOT & 3) == 0 ? 0x3214u : (SLOT & 3) == 1 ? 0x3240u : (SLOT & 3) == 2 ? 0x3410u : 0x4210u;
    R.out[SLOT >> 2] = __byte_perm(R.out[SLOT >> 2], v, on ? ins : 0x3210u);
}
template <int SLOT>
__device__ __forceinline__ uint32_t row_bin(Rc& c, Row& R, uint32_t tt, bool on) {
    rc_renorm(c, on);
    uint32_t ns;
    const uint32_t bit = rc_decide(c, row_get<SLOT>(R), tt, on, ns);
    row_put<SLOT>(R, ns, on);
    return bit;
}
// exponent >= 10 (|value| >= 1024): slots 10 and 31 take several bins each; the row goes through shared memory (rare)
__device__ __forceinline__ int rc_s_wide(Rc& c, uint32_t* out, uint32_t row_a, uint32_t tt_a) {
    uint32_t* rp;
    asm("cvta.shared.u64 %0, %1;" : "=l"(rp) : "l"((unsigned long long)row_a));
    for (int i = 0; i < 8; i++) rp[i] = out[i];
    int e = 10;                                     // slots 1..10 have each answered 1 once
    bool broken = false;
    while (rc_bin(c, row_a + 10, tt_a)) {
        if (++e > 31) { broken = true; break; }
    }
    int a = 0;
    if (!broken) {
        a = 1;
        for (int i = e - 1; i >= 0; i--) a = (a << 1) | (int)rc_bin(c, row_a + 22 + min(i, 9), tt_a);
        if (rc_bin(c, row_a + 21, tt_a)) a = -a;
    } else {
        c.range = 0xFF00; c.end = 0;                // ForceUnderrun (FFV1_RangeCoder.cpp:307-311): the slice is reported broken
    }
    for (int i = 0; i < 8; i++) out[i] = rp[i];
    return a;
}

// mantissa step I (slot 22 + I) for the lanes with I < e
template <int I>
__device__ __forceinline__ void man_step(Rc& c, Row& R, const uint32_t* tq, uint32_t narrow, int e, int& m) {
    const bool mine = narrow && I < e;
    const uint32_t b = row_bin<22 + I>(c, R, tq[11 + I], mine);
    if (mine) m = (m << 1) | (int)b;
}
// rangecoder::s for every lane at once (live: the lane has a sample to decode); returns the signed value
__device__ __forceinline__ int rc_s_row(Rc& c, Row& R, uint32_t row_a, uint32_t tt_a, bool live) {
    // transition pairs of the 20 slots with a fixed place in the walk: slots 0..10 -> tq[0..10], slots 22..30 -> tq[11..19].
    // Each is asked for a few bins before its use, inside the dependent chain of the coder, where the issue slots are free
    uint32_t tq[20];
#define TQ_LOOKUP(SLOT) tq[(SLOT) <= 10 ? (SLOT) : (SLOT) - 11] = lds16(tt_a + 2u * row_get<SLOT>(R))
#define EXP_STEP(K)                                                            \
    going &= row_bin<K>(c, R, tq[K], true); e += (int)going;
    TQ_LOOKUP(0); TQ_LOOKUP(1); TQ_LOOKUP(2); TQ_LOOKUP(3);
    rc_ensure4(c);
    const uint32_t b0 = row_bin<0>(c, R, tq[0], live);
    const uint32_t nz = (live && !b0) ? 1u : 0u;
    uint32_t going = nz;
    int e = 0;
    // exponent bins on slots 1..10: a lane leaves at its first 0. Step k also asks for the pairs of exponent slot k + 3 and of
    // mantissa slot 22 + k - 1 (a symbol that reaches step k has exponent >= k - 1)
    do {
        if (!going) break;
        TQ_LOOKUP(4); TQ_LOOKUP(22); EXP_STEP(1); if (!going) break;
        TQ_LOOKUP(5); TQ_LOOKUP(23); EXP_STEP(2); if (!going) break;
        TQ_LOOKUP(6); TQ_LOOKUP(24); EXP_STEP(3); if (!going) break;
        rc_ensure4(c);
        TQ_LOOKUP(7); TQ_LOOKUP(25); EXP_STEP(4); if (!going) break;
        TQ_LOOKUP(8); TQ_LOOKUP(26); EXP_STEP(5); if (!going) break;
        TQ_LOOKUP(9); TQ_LOOKUP(27); EXP_STEP(6); if (!going) break;
        TQ_LOOKUP(10); TQ_LOOKUP(28); EXP_STEP(7); if (!going) break;
        rc_ensure4(c);
        TQ_LOOKUP(29); EXP_STEP(8); if (!going) break;
        TQ_LOOKUP(30); EXP_STEP(9); if (!going) break;
        EXP_STEP(10);
    } while (0);
#undef EXP_STEP
#undef TQ_LOOKUP
    const uint32_t wide = going;                     // the tenth exponent bin was a 1 as well
    const uint32_t narrow = nz & ~wide;
    // sign state on slot 11 + e (e <= 9): bytes 11..20 of the row; its transition pair is asked for before the mantissa bins
    const uint32_t sslot = 11u + (uint32_t)e;
    const uint32_t swi = sslot >> 2, ssh = (sslot & 3u) * 8u;
    const uint32_t swv = swi == 2 ? R.in[2] : swi == 3 ? R.in[3] : swi == 4 ? R.in[4] : R.in[5];
    const uint32_t sst = (swv >> ssh) & 0xFFu;
    const uint32_t stt = lds16(tt_a + 2u * sst);
    int a = 0;
    if (wide) a = rc_s_wide(c, R.out, row_a, tt_a);
    // mantissa bins on slots 22 + i, i = e - 1 .. 0: every lane walks down from the largest exponent in the warp
    const int emax = __reduce_max_sync(0xffffffffu, narrow ? e : 0);
    int m = 1;
    rc_ensure4(c);
    switch (emax) {
        case 9: man_step<8>(c, R, tq, narrow, e, m);
        case 8: man_step<7>(c, R, tq, narrow, e, m);
        case 7: man_step<6>(c, R, tq, narrow, e, m);
        case 6: man_step<5>(c, R, tq, narrow, e, m);
                rc_ensure4(c);
        case 5: man_step<4>(c, R, tq, narrow, e, m);
        case 4: man_step<3>(c, R, tq, narrow, e, m);
        case 3: man_step<2>(c, R, tq, narrow, e, m);
        case 2: man_step<1>(c, R, tq, narrow, e, m);
                rc_ensure4(c);
        case 1: man_step<0>(c, R, tq, narrow, e, m);
        default: break;
    }
    if (__any_sync(0xffffffffu, narrow)) {
        const bool p = narrow != 0;
        rc_renorm(c, p);
        uint32_t ns;
        const uint32_t bit = rc_decide(c, sst, stt, p, ns);
        if (p) {
            const uint32_t msk = ~(0xFFu << ssh), ins = ns << ssh;
            if (swi == 2) R.out[2] = (R.out[2] & msk) | ins;
            else if (swi == 3) R.out[3] = (R.out[3] & msk) | ins;
            else if (swi == 4) R.out[4] = (R.out[4] & msk) | ins;
            else R.out[5] = (R.out[5] & msk) | ins;
            a = bit ? -m : m;
        }
    }
    return a;
}

// ------------------------------------------------------------------------------------------------------------------
// One decoded slice row (three planes) -> file bytes, all 32 lanes. Inverse of the encoder's load_rgb + forward RCT.
// Returns this lane's count of storage units that differ from the source payload (compare mode).
__device__ __forceinline__ void inv_rct(const DecArgs& A, const int32_t* Y, const int32_t* U, const int32_t* V, int x, int& r, int& g, int& b) {
    const int off = 1 << A.bits;
    g = Y[x]; b = U[x] - off; r = V[x] - off;
    g -= (b + r) >> 2;
    b += g; r += g;
    if (A.swap_bg) { const int t = g; g = b; b = t; }
    const int m = off - 1;
    r &= m; g &= m; b &= m;
}
__device__ __forceinline__ uint32_t diff_bytes16(uint32_t a, uint32_t b) { const uint32_t d = a ^ b; return ((d & 0xFFu) ? 1u : 0u) + ((d & 0xFF00u) ? 1u : 0u); }

__device__ uint32_t pack_row(const DecArgs& A, int lane, int frame, int x0, int yabs, int w,
                             const int32_t* Y, const int32_t* U, const int32_t* V) {
    const size_t rowoff = (size_t)frame * A.frame_bytes + (size_t)yabs * A.row_bytes;
    uint8_t* o = A.out ? A.out + rowoff : nullptr;
    const uint8_t* c = A.cmp ? A.cmp + rowoff : nullptr;
    uint32_t bad = 0;
    const int layout = A.layout;
    if (layout == B200_DPX_RGB_12_PACKED_BE) {
        // component k = 3x + c of the row lives at bit 12k (LSB first) of the row seen as big-endian 32-bit words; a slice
        // owns bits [36 x0, 36 (x0 + w)), so its first and last word may be shared with the neighbouring slices
        const uint32_t b0 = 36u * (uint32_t)x0, b1 = 36u * (uint32_t)(x0 + w);
        const uint32_t w0 = b0 >> 5, w1 = (b1 - 1u) >> 5;
        for (uint32_t wi = w0 + lane; wi <= w1; wi += 32) {
            const uint32_t lo = wi * 32u, hi = lo + 32u;
            const uint32_t klo = max(lo / 12u, b0 / 12u);                 // first component with a bit in this word and in the slice
            const uint32_t khi = min((hi - 1u) / 12u, b1 / 12u - 1u);
            uint32_t val = 0, mask = 0;
            for (uint32_t k = klo; k <= khi; k++) {
                int r, g, b;
                inv_rct(A, Y, U, V, (int)(k / 3u) - x0, r, g, b);
                const uint32_t cc = k % 3u;
                const uint32_t v = (uint32_t)(cc == 0 ? r : cc == 1 ? g : b);
                const int sh = (int)(k * 12u) - (int)lo;                  // -11 .. 31
                if (sh >= 0) { val |= v << sh; mask |= 0xFFFu << sh; }
                else { val |= v >> -sh; mask |= 0xFFFu >> -sh; }
            }
            if (o) {
                uint32_t* po = reinterpret_cast<uint32_t*>(o) + wi;
                if (mask == 0xFFFFFFFFu) *po = bswap32d(val);
                else atomicOr(po, bswap32d(val));                         // the buffer was zeroed before the launch
            }
            if (c) {
                const uint32_t src = bswap32d(reinterpret_cast<const uint32_t*>(c)[wi]);
                if ((src ^ val) & mask) bad++;
            }
        }
        return bad;
    }
    for (int x = lane; x < w; x += 32) {
        int r, g, b;
        inv_rct(A, Y, U, V, x, r, g, b);
        const int xa = x0 + x;
        switch (layout) {
            case B200_DPX_RGB_8: case B200_TIFF_RGB_8: {
                if (o) { uint8_t* q = o + 3 * xa; q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b; }
                if (c) { const uint8_t* q = c + 3 * xa; bad += (q[0] != (uint8_t)r) + (q[1] != (uint8_t)g) + (q[2] != (uint8_t)b); }
                break;
            }
            case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: {
                uint32_t v = ((uint32_t)r << 22) | ((uint32_t)g << 12) | ((uint32_t)b << 2);
                if (layout == B200_DPX_RGB_10_FILLED_A_BE) v = bswap32d(v);
                if (o) reinterpret_cast<uint32_t*>(o)[xa] = v;
                if (c) {
                    const uint32_t d = reinterpret_cast<const uint32_t*>(c)[xa] ^ v;
                    bad += ((d & 0xFFu) != 0) + ((d & 0xFF00u) != 0) + ((d & 0xFF0000u) != 0) + ((d & 0xFF000000u) != 0);
                }
                break;
            }
            default: {      // three 16-bit components; 12-bit Filled A keeps the value in the upper bits
                const bool f12 = layout == B200_DPX_RGB_12_FILLED_A_LE || layout == B200_DPX_RGB_12_FILLED_A_BE;
                const bool be = layout == B200_DPX_RGB_12_FILLED_A_BE || layout == B200_DPX_RGB_16_BE || layout == B200_TIFF_RGB_16_BE;
                uint32_t a0 = (uint32_t)r, a1 = (uint32_t)g, a2 = (uint32_t)b;
                if (f12) { a0 <<= 4; a1 <<= 4; a2 <<= 4; }
                if (be) { a0 = bswap16d(a0); a1 = bswap16d(a1); a2 = bswap16d(a2); }
                if (o) { uint16_t* q = reinterpret_cast<uint16_t*>(o + 6 * (size_t)xa); q[0] = (uint16_t)a0; q[1] = (uint16_t)a1; q[2] = (uint16_t)a2; }
                if (c) {
                    const uint16_t* q = reinterpret_cast<const uint16_t*>(c + 6 * (size_t)xa);
                    bad += diff_bytes16(q[0], a0 & 0xFFFFu) + diff_bytes16(q[1], a1 & 0xFFFFu) + diff_bytes16(q[2], a2 & 0xFFFFu);
                }
            }
        }
    }
    return bad;
}

// ------------------------------------------------------------------------------------------------------------------
// k_decode
struct DecSmem {
    int16_t* qtab; uint8_t* trans; uint32_t* crc; uint8_t* rows;
};
__host__ __device__ inline size_t dec_smem_bytes(int nsets, int) {
    return (size_t)nsets * 5 * 256 * 2 + 512 + 1024 + (size_t)kDecWarpsPerCta * 32 * kRowStride;
}

__global__ void __launch_bounds__(32 * kDecWarpsPerCta, B200_DEC_MIN_CTAS) k_decode(const __grid_constant__ DecArgs A, int nframes) {
    extern __shared__ __align__(16) uint8_t smem[];
    DecSmem S;
    S.qtab = reinterpret_cast<int16_t*>(smem);
    S.trans = smem + (size_t)A.nsets * 5 * 256 * 2;
    S.crc = reinterpret_cast<uint32_t*>(S.trans + 512);
    S.rows = reinterpret_cast<uint8_t*>(S.crc + 256);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < A.nsets * 5 * 256; i += blockDim.x) S.qtab[i] = A.qtab[i];
    for (int i = tid; i < 256; i += blockDim.x) reinterpret_cast<uint16_t*>(S.trans)[i] = (uint16_t)(A.trans[i] | (A.trans[256 + i] << 8));
    for (int i = tid; i < 256; i += blockDim.x) S.crc[i] = A.crc_table[i];
    __syncthreads();
    const uint32_t tt_a = smem_a(S.trans), qtab_a = smem_a(S.qtab);   // tt: (next state after 0) | (next state after 1) << 8
    const uint32_t row_a = smem_a(S.rows) + (uint32_t)((warp * 32 + lane) * kRowStride);
    const int spw = A.spw;
    const int total = nframes * A.nslices;
    const int wid = blockIdx.x * kDecWarpsPerCta + warp;
    const size_t line_stride = (size_t)9 * A.wpad;            // ints per slice: [plane][row % 3][wpad]
    const size_t state_stride = (size_t)A.maxctx * 32;        // bytes per (slice, plane-set)

    // ---- prologue, all lanes, one slice after the other: CRC of the slice (FFV1_Slice.cpp:247-249), rows above the slice = 0
    for (int j = 0; j < spw; j++) {
        const int sidx = wid * spw + j;
        if (sidx >= total) break;
        const uint32_t size = A.sl_size[sidx];
        int32_t* ln = A.lines + (size_t)sidx * line_stride;
        for (int i = lane; i < (int)line_stride; i += 32) ln[i] = 0;
        if (!size || !A.ec) continue;
        const uint8_t* p = A.packets + A.sl_off[sidx];
        const uint32_t L = (size + 31u) / 32u;
        const uint32_t b0 = min(size, lane * L), b1 = min(size, b0 + L);
        uint32_t crc = 0;
        for (uint32_t i = b0; i < b1; i++) crc = (crc << 8) ^ S.crc[(crc >> 24) ^ p[i]];
        // times x^(8 * bytes after this lane's piece) in GF(2)[x]/P: the CRC is linear (see k_pack)
        uint32_t after = size - b1, pw = 0x100u;
        while (after) {
            if (after & 1u) {
                uint32_t r = 0;
                for (int i = 31; i >= 0; i--) { r = (r << 1) ^ ((r >> 31) ? 0x04C11DB7u : 0u); if ((pw >> i) & 1u) r ^= crc; }
                crc = r;
            }
            after >>= 1;
            if (after) {
                uint32_t r = 0;
                for (int i = 31; i >= 0; i--) { r = (r << 1) ^ ((r >> 31) ? 0x04C11DB7u : 0u); if ((pw >> i) & 1u) r ^= pw; }
                pw = r;
            }
        }
        for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
        if (lane == 0 && crc) atomicOr(&A.status[sidx / A.nslices], (uint32_t)B200_DEC_BAD_CRC);
    }
    __syncwarp();

    // ---- per-lane slice state
    const int sidx = wid * spw + lane;
    bool active = lane < spw && sidx < total && A.sl_size[sidx < total ? sidx : 0] != 0;
    const int frame = sidx < total ? sidx / A.nslices : 0;
    Rc rc;
    rc_dead(rc);
    bool hdr_broken = false;
    const uint8_t* slice_p = nullptr;
    int gx0 = 0, gy0 = 0, gw = 0, gh = 0;
    int set0 = 0, set1 = 0;                     // quantisation-table set of plane-set 0 (Y) and 1 (Cb, Cr)
    uint32_t slice_bytes = 0;
    uint32_t myflags = 0;
    if (active) {
        const uint8_t* p = A.packets + A.sl_off[sidx];
        slice_bytes = A.sl_size[sidx];
        slice_p = p;
        rc_init(rc, p, slice_bytes - (uint32_t)A.tail);
        for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(S.rows + (size_t)(warp * 32 + lane) * kRowStride)[i] = 0x80808080u;
        if (A.sl_off[sidx] == A.pkt_off[frame]) {               // first slice of the packet: the keyframe bin (FFV1_Slice.cpp:221-225)
            const uint32_t key = rc_bin(rc, row_a, tt_a);
            sts8(row_a, 128u);
            if (!key) { myflags |= B200_DEC_BAD_HEADER; active = false; }
        }
        if (active) {
            // slice header (FFV1_Slice.cpp:113-177): all symbols on one set of 32 states
            // slice_x, slice_y, slice_width - 1, slice_height - 1, quant_table_set_index x 2, picture_structure, sar_num, sar_den
            uint32_t hv[9];
#pragma unroll 1
            for (int k = 0; k < 9; k++) hv[k] = rc_u(rc, row_a, tt_a, hdr_broken);
            const uint32_t sx = hv[0], sy = hv[1], x2 = sx + hv[2] + 1, y2 = sy + hv[3] + 1, q0 = hv[4], q1 = hv[5];
            if (sx >= (uint32_t)A.num_h || sy >= (uint32_t)A.num_h || sy >= (uint32_t)A.num_v || x2 > (uint32_t)A.num_h || y2 > (uint32_t)A.num_v ||
                q0 >= (uint32_t)A.nsets || q1 >= (uint32_t)A.nsets || hdr_broken || rc_consumed(rc) > rc.end + 1u) {
                myflags |= B200_DEC_BAD_HEADER; active = false;
            } else {
                gx0 = (int)((uint64_t)sx * A.W / A.num_h);
                gy0 = (int)((uint64_t)sy * A.H / A.num_v);
                gw = (int)((uint64_t)x2 * A.W / A.num_h) - gx0;
                gh = (int)((uint64_t)y2 * A.H / A.num_v) - gy0;
                set0 = (int)q0; set1 = (int)q1;
                if (gw <= 0 || gh <= 0 || gw > A.wpad) { myflags |= B200_DEC_BAD_HEADER; active = false; }
            }
        }
    }

    int32_t* lines = A.lines + (size_t)(sidx < total ? sidx : 0) * line_stride;
    uint8_t* states = A.states + (size_t)(sidx < total ? sidx : 0) * 2 * state_stride;
    const int bits_mask = (1 << A.bits_max) - 1;
    int x = 0, y = 0, pl = 0;
    int L = 0, LT = 0, T = 0, LL = 0, RT = 0, TT = 0;
    int cached = -1;                            // (plane-set << 16 | context) whose 32 states are in R
    Row R;
    for (int i = 0; i < 8; i++) R.in[i] = R.out[i] = 0;
    const int32_t *prv = nullptr, *pp2 = nullptr;
    int32_t* cur = nullptr;
    uint32_t qa = 0;                            // shared address of the five tables of the current plane's set
    bool is5 = false;
    bool fresh = true;                          // (x == 0): row pointers and border values must be set up
    bool rowdone = false;

    if (!active) rc_dead(rc);
    while (__any_sync(0xffffffffu, active)) {
        // ---- phase A (lanes with a slice): neighbours, context, state row of the context
        int pred = 0, nRT = 0, nTT = 0;
        bool neg = false;
        if (active) {
            if (fresh) {
                const int ps = pl ? 1 : 0;
                const int set = ps ? set1 : set0;
                qa = qtab_a + (uint32_t)set * 5u * 256u * 2u;
                is5 = S.qtab[(size_t)set * 5 * 256 + 3 * 256 + 127] != 0;
                int32_t* base = lines + (size_t)pl * 3 * A.wpad;
                cur = base + (size_t)(y % 3) * A.wpad;
                prv = base + (size_t)((y + 2) % 3) * A.wpad;
                pp2 = base + (size_t)((y + 1) % 3) * A.wpad;
                // borders (FFV1_Slice.cpp:430-433): sample[-1] of this row = first sample of the row above, the row above
                // carries what its own [-1] was (the first sample of the row above it)
                T = prv[0]; L = T; LT = pp2[0]; LL = 0;
                RT = prv[gw > 1 ? 1 : 0];
                TT = pp2[0];
                fresh = false;
            }
            // neighbours of the next sample that do not depend on this one, asked for now
            nRT = prv[min(x + 2, gw - 1)];
            nTT = pp2[min(x + 1, gw - 1)];
            int ctx = lds_s16(qa + (uint32_t)((L - LT) & 255) * 2u) + lds_s16(qa + 512u + (uint32_t)((LT - T) & 255) * 2u) +
                      lds_s16(qa + 1024u + (uint32_t)((T - RT) & 255) * 2u);
            if (is5) ctx += lds_s16(qa + 1536u + (uint32_t)((LL - L) & 255) * 2u) + lds_s16(qa + 2048u + (uint32_t)((TT - T) & 255) * 2u);
            pred = median3d(L, L + T - LT, T);
            neg = ctx < 0;
            if (neg) ctx = -ctx;
            const int want = ((pl ? 1 : 0) << 16) | ctx;
            if (want != cached) {
                // the row in R goes back to the table in global memory, the wanted one comes from there. (A direct-mapped cache
                // of rows in shared memory in front of the table was measured: on noisy 16-bit material the quantised context is
                // spread over all 10 126 rows, 94 % of the look-ups missed and the bookkeeping cost more than it saved.)
                if (cached >= 0) {
                    uint4* g = reinterpret_cast<uint4*>(states + (size_t)(cached >> 16) * state_stride + (size_t)(cached & 0xFFFF) * 32);
                    g[0] = make_uint4(R.out[0], R.out[1], R.out[2], R.out[3]);
                    g[1] = make_uint4(R.out[4], R.out[5], R.out[6], R.out[7]);
                }
                const uint4* g = reinterpret_cast<const uint4*>(states + (size_t)(want >> 16) * state_stride + (size_t)ctx * 32);
                const uint4 a = g[0], b = g[1];
                R.out[0] = a.x; R.out[1] = a.y; R.out[2] = a.z; R.out[3] = a.w; R.out[4] = b.x; R.out[5] = b.y; R.out[6] = b.z; R.out[7] = b.w;
                cached = want;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) R.in[i] = R.out[i];
        // ---- phase B (all lanes, one instruction stream): the symbol
        const int d = rc_s_row(rc, R, row_a, tt_a, active);
        // ---- phase C: the sample
        if (active) {
            const int val = (pred + (neg ? -d : d)) & bits_mask;
            cur[x] = val;
            LL = L; L = val; LT = T; T = RT; RT = nRT; TT = nTT;
            if (++x == gw) {
                x = 0; fresh = true;
                if (++pl == 3) { pl = 0; rowdone = true; }
            }
        }
        uint32_t ready = __ballot_sync(0xffffffffu, rowdone);
        if (ready) {
            __syncwarp();
            while (ready) {
                const int src = __ffs(ready) - 1;
                ready &= ready - 1;
                const int f = __shfl_sync(0xffffffffu, frame, src);
                const int px0 = __shfl_sync(0xffffffffu, gx0, src), py = __shfl_sync(0xffffffffu, gy0 + y, src), pw = __shfl_sync(0xffffffffu, gw, src);
                const int yy = __shfl_sync(0xffffffffu, y, src);
                const int32_t* lb = A.lines + (size_t)(wid * spw + src) * line_stride + (size_t)(yy % 3) * A.wpad;
                uint32_t bad = pack_row(A, lane, f, px0, py, pw, lb, lb + (size_t)3 * A.wpad, lb + (size_t)6 * A.wpad);
                if (A.cmp) {
                    for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
                    if (lane == 0 && bad) atomicAdd(&A.mismatch[f], (unsigned long long)bad);
                }
            }
            __syncwarp();
            if (rowdone) {
                rowdone = false;
                if (++y == gh) {
                    // end of the slice: terminator bin (FFV1_Slice.cpp:335-341), underrun / junk / error_status checks
                    // (rangecoder::BytesUsed / IsUnderrun, FFV1_RangeCoder.cpp:51-62)
                    cached = -1;
                    sts8(row_a, 129u);
                    rc_bin(rc, row_a, tt_a);
                    const uint32_t adj = rc.range < 0x100u ? 0u : 1u;
                    const uint32_t pos = rc_consumed(rc);
                    if (rc.end == 0 || pos - adj > rc.end) myflags |= B200_DEC_UNDERRUN;
                    const uint32_t used = pos > rc.end ? rc.end : pos - adj;
                    if (used < rc.end) myflags |= B200_DEC_JUNK;
                    if (A.ec && slice_p[slice_bytes - 5]) myflags |= B200_DEC_ERROR_STATUS;
                    atomicAdd(&A.counters[0], 1ull);
                    atomicAdd(&A.counters[1], (unsigned long long)gw * gh * 3ull);
                    active = false;
                    rc_dead(rc);
                }
            }
        }
    }
    if (myflags) atomicOr(&A.status[frame], myflags);
}

}  // namespace

cudaError_t launch_dec_index(const DecArgs& a, int nframes, cudaStream_t s) {
    k_dec_index<<<(nframes + 127) / 128, 128, 0, s>>>(a, nframes);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecArgs& a, int nframes, cudaStream_t s) {
    const size_t smem = dec_smem_bytes(a.nsets, a.spw);
    auto fn = k_decode;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int total = nframes * a.nslices;
    const int warps = (total + a.spw - 1) / a.spw;
    fn<<<(warps + kDecWarpsPerCta - 1) / kDecWarpsPerCta, 32 * kDecWarpsPerCta, smem, s>>>(a, nframes);
    return cudaGetLastError();
}

}  // namespace b200
