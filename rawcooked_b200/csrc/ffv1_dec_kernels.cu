// sm_100a kernels of the B200 FFV1 decoder (`--check` side; interface in ffv1_dec.h, C ABI in include/b200dec.h).
//
// The decoder is serial per slice in both the context model and the range coder: every decoded sample is the left
// neighbour of the next one's predictor and context. The parallelism is therefore slices x frames in flight: every slice
// of the batch is decoded by ONE lane, `spw` slices share a warp (their lanes step through their streams sample by sample in
// lockstep, so the warp issues one instruction stream for spw slices), and a finished row of any of them is turned into
// file bytes by all 32 lanes (inverse RCT, byte layout, store and/or compare with the source payload).
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
// renormalisation needs is fetched when the next bin is asked for, as the reference does, so that the byte accounting at the
// end of the slice (BytesUsed) is the reference's
struct Rc {
    const uint8_t* buf;
    uint32_t low, range, pos, end;
    bool underrun;
};
__device__ __forceinline__ void rc_init(Rc& c, const uint8_t* buf, uint32_t size) {
    c.buf = buf; c.end = size; c.low = size ? buf[0] : 0; c.range = 0xFF; c.pos = 1; c.underrun = false;
}
// one bin on the adaptive state at shared-memory address sa; trans_a: [0..255] next state after 0, [256..511] after 1
__device__ __forceinline__ uint32_t rc_bin(Rc& c, uint32_t sa, uint32_t trans_a) {
    const uint32_t st = lds8(sa);
    if (c.range < 0x100u) {
        c.low <<= 8;
        if (c.pos > c.end) { c.underrun = true; return 0; }
        if (c.pos < c.end) c.low |= c.buf[c.pos];
        c.range <<= 8;
        c.pos++;
    }
    const uint32_t r1 = (c.range * st) >> 8;
    c.range -= r1;
    const uint32_t bit = c.low >= c.range ? 1u : 0u;
    if (bit) { c.low -= c.range; c.range = r1; }
    sts8(sa, lds8(trans_a + st + (bit << 8)));
    return bit;
}
// rangecoder::u (FFV1_RangeCoder.cpp:105-132) on the 32 states at row_a
__device__ __forceinline__ uint32_t rc_u(Rc& c, uint32_t row_a, uint32_t trans_a) {
    if (rc_bin(c, row_a, trans_a)) return 0;
    int e = 0;
    while (rc_bin(c, row_a + 1 + min(e, 9), trans_a)) {
        if (++e > 31) { c.underrun = true; c.range = 0; return 0; }
    }
    uint32_t a = 1;
    for (int i = e - 1; i >= 0; i--) a = (a << 1) | rc_bin(c, row_a + 22 + min(i, 9), trans_a);
    return a;
}
// rangecoder::s (FFV1_RangeCoder.cpp:135-171, the rolled form)
__device__ __forceinline__ int rc_s(Rc& c, uint32_t row_a, uint32_t trans_a) {
    if (rc_bin(c, row_a, trans_a)) return 0;
    int e = 0;
    while (rc_bin(c, row_a + 1 + min(e, 9), trans_a)) {
        if (++e > 31) { c.underrun = true; c.range = 0; return 0; }
    }
    int a = 1;
    for (int i = e - 1; i >= 0; i--) a = (a << 1) | (int)rc_bin(c, row_a + 22 + min(i, 9), trans_a);
    return rc_bin(c, row_a + 11 + min(e, 10), trans_a) ? -a : a;
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
__host__ __device__ inline size_t dec_smem_bytes(int nsets) {
    return (size_t)nsets * 5 * 256 * 2 + 512 + 1024 + (size_t)kDecWarpsPerCta * 32 * kRowStride;
}

__global__ void __launch_bounds__(32 * kDecWarpsPerCta) k_decode(const __grid_constant__ DecArgs A, int nframes) {
    extern __shared__ __align__(16) uint8_t smem[];
    DecSmem S;
    S.qtab = reinterpret_cast<int16_t*>(smem);
    S.trans = smem + (size_t)A.nsets * 5 * 256 * 2;
    S.crc = reinterpret_cast<uint32_t*>(S.trans + 512);
    S.rows = reinterpret_cast<uint8_t*>(S.crc + 256);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < A.nsets * 5 * 256; i += blockDim.x) S.qtab[i] = A.qtab[i];
    for (int i = tid; i < 512; i += blockDim.x) S.trans[i] = A.trans[i];
    for (int i = tid; i < 256; i += blockDim.x) S.crc[i] = A.crc_table[i];
    __syncthreads();
    const uint32_t trans_a = smem_a(S.trans), qtab_a = smem_a(S.qtab);
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
    rc.buf = nullptr; rc.low = rc.range = rc.pos = rc.end = 0; rc.underrun = false;
    int gx0 = 0, gy0 = 0, gw = 0, gh = 0;
    int set0 = 0, set1 = 0;                     // quantisation-table set of plane-set 0 (Y) and 1 (Cb, Cr)
    uint32_t slice_bytes = 0;
    uint32_t myflags = 0;
    if (active) {
        const uint8_t* p = A.packets + A.sl_off[sidx];
        slice_bytes = A.sl_size[sidx];
        rc_init(rc, p, slice_bytes - (uint32_t)A.tail);
        for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(S.rows + (size_t)(warp * 32 + lane) * kRowStride)[i] = 0x80808080u;
        if (A.sl_off[sidx] == A.pkt_off[frame]) {               // first slice of the packet: the keyframe bin (FFV1_Slice.cpp:221-225)
            const uint32_t key = rc_bin(rc, row_a, trans_a);
            sts8(row_a, 128u);
            if (!key) { myflags |= B200_DEC_BAD_HEADER; active = false; }
        }
        if (active) {
            // slice header (FFV1_Slice.cpp:113-177): all symbols on one set of 32 states
            const uint32_t sx = rc_u(rc, row_a, trans_a), sy = rc_u(rc, row_a, trans_a);
            const uint32_t sw1 = rc_u(rc, row_a, trans_a), sh1 = rc_u(rc, row_a, trans_a);
            const uint32_t x2 = sx + sw1 + 1, y2 = sy + sh1 + 1;
            const uint32_t q0 = rc_u(rc, row_a, trans_a), q1 = rc_u(rc, row_a, trans_a);
            rc_u(rc, row_a, trans_a);           // picture_structure
            rc_u(rc, row_a, trans_a);           // sar_num
            rc_u(rc, row_a, trans_a);           // sar_den
            if (sx >= (uint32_t)A.num_h || sy >= (uint32_t)A.num_h || sy >= (uint32_t)A.num_v || x2 > (uint32_t)A.num_h || y2 > (uint32_t)A.num_v ||
                q0 >= (uint32_t)A.nsets || q1 >= (uint32_t)A.nsets || rc.underrun) {
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
    int cached = -1;                            // (plane-set << 16 | context) whose 32 states are staged at row_a
    const int32_t *prv = nullptr, *pp2 = nullptr;
    int32_t* cur = nullptr;
    uint32_t qa = 0;                            // shared address of the five tables of the current plane's set
    bool is5 = false;
    bool fresh = true;                          // (x == 0): row pointers and border values must be set up
    bool rowdone = false;

    auto write_back = [&]() {
        if (cached >= 0) {
            const uint4* r = reinterpret_cast<const uint4*>(S.rows + (size_t)(warp * 32 + lane) * kRowStride);
            uint4* g = reinterpret_cast<uint4*>(states + (size_t)(cached >> 16) * state_stride + (size_t)(cached & 0xFFFF) * 32);
            g[0] = r[0]; g[1] = r[1];
        }
    };

    while (__any_sync(0xffffffffu, active)) {
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
            const int nRT = prv[min(x + 2, gw - 1)];
            const int nTT = pp2[min(x + 1, gw - 1)];
            int ctx = lds_s16(qa + (uint32_t)((L - LT) & 255) * 2u) + lds_s16(qa + 512u + (uint32_t)((LT - T) & 255) * 2u) +
                      lds_s16(qa + 1024u + (uint32_t)((T - RT) & 255) * 2u);
            if (is5) ctx += lds_s16(qa + 1536u + (uint32_t)((LL - L) & 255) * 2u) + lds_s16(qa + 2048u + (uint32_t)((TT - T) & 255) * 2u);
            const int pred = median3d(L, L + T - LT, T);
            const bool neg = ctx < 0;
            if (neg) ctx = -ctx;
            const int want = ((pl ? 1 : 0) << 16) | ctx;
            if (want != cached) {
                write_back();
                const uint4* g = reinterpret_cast<const uint4*>(states + (size_t)(want >> 16) * state_stride + (size_t)ctx * 32);
                uint4* r = reinterpret_cast<uint4*>(S.rows + (size_t)(warp * 32 + lane) * kRowStride);
                const uint4 a = g[0], b = g[1];
                r[0] = a; r[1] = b;
                cached = want;
            }
            const int d = rc_s(rc, row_a, trans_a);
            const int val = (pred + (neg ? -d : d)) & bits_mask;
            cur[x] = val;
            LL = L; L = val; LT = T; T = RT; RT = nRT; TT = nTT;
            if (++x == gw) {
                x = 0; fresh = true;
                if (++pl == 3) { pl = 0; rowdone = true; }
            }
        }
        __syncwarp();
        uint32_t ready = __ballot_sync(0xffffffffu, rowdone);
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
                write_back();
                cached = -1;
                sts8(row_a, 129u);
                rc_bin(rc, row_a, trans_a);
                const uint32_t adj = rc.range < 0x100u ? 0u : 1u;
                if (rc.underrun || rc.pos - adj > rc.end) myflags |= B200_DEC_UNDERRUN;
                const uint32_t used = rc.pos > rc.end ? rc.end : rc.pos - adj;
                if (used < rc.end) myflags |= B200_DEC_JUNK;
                if (A.ec && rc.buf[rc.end + 3]) myflags |= B200_DEC_ERROR_STATUS;
                atomicAdd(&A.counters[0], 1ull);
                atomicAdd(&A.counters[1], (unsigned long long)gw * gh * 3ull);
                active = false;
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
    const size_t smem = dec_smem_bytes(a.nsets);
    cudaError_t e = cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int total = nframes * a.nslices;
    const int warps = (total + a.spw - 1) / a.spw;
    k_decode<<<(warps + kDecWarpsPerCta - 1) / kDecWarpsPerCta, 32 * kDecWarpsPerCta, smem, s>>>(a, nframes);
    return cudaGetLastError();
}

}  // namespace b200
