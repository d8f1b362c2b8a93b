// Once-per-stream host work of the B200 FFV1 encoder (see ffv1_host.h).
#include "ffv1_host.h"

#include <cstring>

#include "../../include/b200enc.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// layouts
int layout_bits(int layout) {
    switch (layout) {
        case B200_DPX_RGB_8: case B200_TIFF_RGB_8: return 8;
        case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: return 10;
        case B200_DPX_RGB_12_FILLED_A_LE: case B200_DPX_RGB_12_FILLED_A_BE: case B200_DPX_RGB_12_PACKED_BE: return 12;
        case B200_DPX_RGB_16_LE: case B200_DPX_RGB_16_BE: case B200_TIFF_RGB_16_LE: case B200_TIFF_RGB_16_BE: return 16;
    }
    return 0;
}

// DPX image rows are padded to 32 bits (reference: Source/Lib/Uncompressed/DPX/DPX.cpp:478-482); TIFF strips are not
size_t layout_row_bytes(uint32_t w, int layout) {
    switch (layout) {
        case B200_DPX_RGB_8: return ((size_t)w * 3 + 3) & ~(size_t)3;
        case B200_TIFF_RGB_8: return (size_t)w * 3;
        case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: return (size_t)w * 4;
        case B200_DPX_RGB_12_PACKED_BE: return (((size_t)w * 36 + 31) / 32) * 4;
        // 16-bit DPX is not a "Filled" packing: its rows take the 32-bit line padding of DPX.cpp:478-482 (odd widths: + 2 bytes),
        // 12-bit Filled-A rows are whole 6-byte blocks (DPX.cpp:463-475), TIFF strips are tight
        case B200_DPX_RGB_16_LE: case B200_DPX_RGB_16_BE: return ((size_t)w * 6 + 3) & ~(size_t)3;
        case B200_DPX_RGB_12_FILLED_A_LE: case B200_DPX_RGB_12_FILLED_A_BE:
        case B200_TIFF_RGB_16_LE: case B200_TIFF_RGB_16_BE: return (size_t)w * 6;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CRC-32/MPEG-2 polynomial, init 0, no reflection, no final xor: the parity the reference checks with
// ZenCRC32(...) == 0 (Source/Lib/Utils/CRC32/ZenCRC32.cpp:1097-1135; FFV1_Frame.cpp:114-117, FFV1_Slice.cpp:247-249)
namespace {
struct CrcTable {
    uint32_t t[256];
    CrcTable() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
            t[i] = c;
        }
    }
};
const CrcTable g_crc;
}  // namespace
const uint32_t* crc32_mpeg_table() { return g_crc.t; }
uint32_t crc32_mpeg(const uint8_t* d, size_t n, uint32_t crc) {
    for (size_t i = 0; i < n; i++) crc = (crc << 8) ^ g_crc.t[(crc >> 24) ^ d[i]];
    return crc;
}

// ------------------------------------------------------------------------------------------------
// `-slices N` -> grid, the search ffmpeg's encoder does (v from 1 or 2, h in [v, 2v]); the set of counts it
// accepts is what the reference's test/slices.sh:12 pins
int slice_grid(uint32_t w, uint32_t h, int slices, int bits, int* num_h, int* num_v) {
    int v = (w > 352 || h > 288 || !slices) ? 2 : 1;
    if ((uint32_t)v > h) v = (int)h;
    for (; v <= 32; v++)
        for (int hh = v; hh <= 2 * v; hh++) {
            if ((uint32_t)hh > w || (uint32_t)v > h) continue;
            int64_t maxw = (w + hh - 1) / hh, maxh = (h + v - 1) / v;
            if (maxw * maxh * (bits + 1) * 3 > (8 << 24)) continue;
            if (slices == hh * v && slices <= 1024) { *num_h = hh; *num_v = v; return 0; }
            if (maxw * maxh > 360 * 288) continue;
            if (!slices) { *num_h = hh; *num_v = v; return 0; }
        }
    return B200_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// A byte-oriented binary range encoder used only for the few hundred bins of the ConfigurationRecord, and as
// a "recorder" for the slice-header bins that the GPU coder replays.
namespace {
class HostRangeEncoder {
  public:
    explicit HostRangeEncoder(const uint8_t* one) {
        std::memcpy(one_, one, 256);
        zero_[0] = 0;
        for (int i = 1; i < 256; i++) zero_[i] = (uint8_t)(256 - one_[256 - i]);
    }
    std::vector<uint16_t> recorded;   // state | bit << 8 of every bin, in order

    void bin(uint8_t& state, bool bit) {
        recorded.push_back((uint16_t)(state | (bit ? 256 : 0)));
        uint32_t r1 = (range_ * state) >> 8;
        if (bit) { low_ += range_ - r1; range_ = r1; state = one_[state]; }
        else     { range_ -= r1; state = zero_[state]; }
        shift();
    }
    // unsigned / signed symbol: unary exponent, mantissa, sign (decoder: rangecoder::u / ::s, FFV1_RangeCoder.cpp:105-171)
    void symbol(uint8_t* st, int v, bool is_signed) {
        if (v == 0) { bin(st[0], true); return; }
        unsigned a = v < 0 ? (unsigned)-v : (unsigned)v;
        int e = 0;
        while ((a >> (e + 1)) != 0) e++;
        bin(st[0], false);
        for (int i = 0; i < e; i++) bin(st[1 + (i < 9 ? i : 9)], true);
        bin(st[1 + (e < 9 ? e : 9)], false);
        for (int i = e - 1; i >= 0; i--) bin(st[22 + (i < 9 ? i : 9)], (a >> i) & 1);
        if (is_signed) bin(st[11 + (e < 10 ? e : 10)], v < 0);
    }
    std::vector<uint8_t> finish() {
        range_ = 0xFF; low_ += 0xFF; shift();
        range_ = 0xFF; shift();
        return bytes_;
    }

  private:
    void shift() {
        while (range_ < 0x100) {
            if (pending_ < 0) pending_ = (int)(low_ >> 8);
            else if (low_ <= 0xFF00) { flush(pending_, 0xFF); pending_ = (int)(low_ >> 8); }
            else if (low_ >= 0x10000) { flush(pending_ + 1, 0x00); pending_ = (int)((low_ >> 8) & 0xFF); }
            else run_++;
            low_ = (low_ & 0xFF) << 8;
            range_ <<= 8;
        }
    }
    void flush(int first, int fill) {
        bytes_.push_back((uint8_t)first);
        for (; run_; run_--) bytes_.push_back((uint8_t)fill);
    }
    uint8_t one_[256], zero_[256];
    uint32_t low_ = 0, range_ = 0xFF00;
    int pending_ = -1;
    size_t run_ = 0;
    std::vector<uint8_t> bytes_;
};

}  // namespace
// default transition table of the FFV1 range coder (the reference holds it as a literal,
// Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:35-55; generated here from its defining recurrence, p += (1-p)*0.05)
void default_one_state(uint8_t one[256]) {
    const int64_t unit = (int64_t)1 << 32;
    const int64_t factor = (int64_t)(0.05 * 4294967296.0);
    const int cap = 248;
    std::memset(one, 0, 256);
    int64_t p = unit / 2;
    int prev = 0;
    for (int i = 0; i < 128; i++) {
        int p8 = (int)((256 * p + unit / 2) >> 32);
        if (p8 <= prev) p8 = prev + 1;
        if (prev && prev < 256 && p8 <= cap) one[prev] = (uint8_t)p8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        prev = p8;
    }
    for (int i = 256 - cap; i <= cap; i++) {
        if (one[i]) continue;
        p = ((int64_t)i * unit + 128) >> 8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        int p8 = (int)((256 * p + unit / 2) >> 32);
        if (p8 <= i) p8 = i + 1;
        if (p8 > cap) p8 = cap;
        one[i] = (uint8_t)p8;
    }
}
namespace {

// transition table sent on the wire for `-coder 1` (coder_type 2): the table ffmpeg sends, so that packets are
// byte-identical to the reference pipeline's (read back by parameters::Parse, FFV1_Parameters.cpp:41-55)
const uint8_t kCustomOneState[256] = {
      0, 10, 10, 10, 10, 16, 16, 16, 28, 16, 16, 29, 42, 49, 20, 49, 59, 25, 26, 26, 27, 31, 33, 33, 33, 34, 34, 37, 67, 38, 39, 39,
     40, 40, 41, 79, 43, 44, 45, 45, 48, 48, 64, 50, 51, 52, 88, 52, 53, 74, 55, 57, 58, 58, 74, 60,101, 61, 62, 84, 66, 66, 68, 69,
     87, 82, 71, 97, 73, 73, 82, 75,111, 77, 94, 78, 87, 81, 83, 97, 85, 83, 94, 86, 99, 89, 90, 99,111, 92, 93,134, 95, 98,105, 98,
    105,110,102,108,102,118,103,106,106,113,109,112,114,112,116,125,115,116,117,117,126,119,125,121,121,123,145,124,126,131,127,129,
    165,130,132,138,133,135,145,136,137,139,146,141,143,142,144,148,147,155,151,149,151,150,152,157,153,154,156,168,158,162,161,160,
    172,163,169,164,166,184,167,170,177,174,171,173,182,176,180,178,175,189,179,181,186,183,192,185,200,187,191,188,190,197,193,196,
    197,194,195,196,198,202,199,201,210,203,207,204,205,206,208,214,209,211,221,212,213,215,224,216,217,218,219,220,222,228,223,225,
    226,224,227,229,240,230,231,232,233,234,235,236,238,239,237,242,241,243,242,244,245,246,247,248,249,250,251,252,252,253,254,255,
};

// quantisation tables: thresholds of the first 128 entries, negatives mirrored exactly as the decoder rebuilds them
// (parameters::QuantizationTable, FFV1_Parameters.cpp:222-253)
void make_qtab(int16_t* t, const int* thr, int n, int scale) {
    for (int i = 0; i < 128; i++) {
        int level = 0;
        while (level < n && i >= thr[level]) level++;
        t[i] = (int16_t)(level * scale);
    }
    for (int i = 1; i < 128; i++) t[256 - i] = (int16_t)(-t[i]);
    t[128] = (int16_t)(-t[127]);
}
struct QuantSets { int16_t q[2][5][256]; int nctx[2]; };
void make_quant_sets(int bits, QuantSets* Q) {
    std::memset(Q, 0, sizeof(*Q));
    if (bits <= 8) {
        const int a[] = {1, 2, 5, 12, 35}, b[] = {1, 4};
        make_qtab(Q->q[0][0], a, 5, 1); make_qtab(Q->q[0][1], a, 5, 11); make_qtab(Q->q[0][2], a, 5, 121);
        make_qtab(Q->q[1][0], a, 5, 1); make_qtab(Q->q[1][1], a, 5, 11);
        make_qtab(Q->q[1][2], b, 2, 121); make_qtab(Q->q[1][3], b, 2, 605); make_qtab(Q->q[1][4], b, 2, 3025);
        Q->nctx[0] = (11 * 11 * 11 + 1) / 2; Q->nctx[1] = (11 * 11 * 125 + 1) / 2;
    } else {
        const int a[] = {5, 13, 27, 56}, b[] = {11, 50};
        make_qtab(Q->q[0][0], a, 4, 1); make_qtab(Q->q[0][1], a, 4, 9); make_qtab(Q->q[0][2], a, 4, 81);
        make_qtab(Q->q[1][0], a, 4, 1); make_qtab(Q->q[1][1], a, 4, 9);
        make_qtab(Q->q[1][2], b, 2, 81); make_qtab(Q->q[1][3], b, 2, 405); make_qtab(Q->q[1][4], b, 2, 2025);
        Q->nctx[0] = (9 * 9 * 9 + 1) / 2; Q->nctx[1] = (9 * 9 * 125 + 1) / 2;
    }
}
void put_qtab(HostRangeEncoder& rc, const int16_t* t) {
    uint8_t st[32];
    std::memset(st, 128, sizeof st);
    int run_start = 0;
    for (int i = 1; i < 128; i++)
        if (t[i] != t[i - 1]) { rc.symbol(st, i - run_start - 1, false); run_start = i; }
    rc.symbol(st, 128 - run_start - 1, false);
}
}  // namespace

int build_stream(uint32_t width, uint32_t height, int layout, int slices, int context, int ec, Ffv1Stream* S, const char** err) {
    static const char* e_layout = "unsupported layout";
    static const char* e_dim = "bad dimensions";
    static const char* e_grid = "no slice grid for this -slices value";
    S->bits = layout_bits(layout);
    if (!S->bits) { *err = e_layout; return B200_ERR_INVALID; }
    if (!width || !height || width > 65535 || height > 65535) { *err = e_dim; return B200_ERR_INVALID; }
    S->width = width; S->height = height; S->layout = layout;
    S->sbits = S->bits <= 8 ? 9 : S->bits + 1;
    S->swap_bg = S->bits > 8 && S->bits < 16;
    S->context = context ? 1 : 0;
    S->ec = ec ? 1 : 0;
    S->row_bytes = layout_row_bytes(width, layout);
    S->frame_bytes = S->row_bytes * height;
    if (slice_grid(width, height, slices, S->bits, &S->num_h, &S->num_v)) { *err = e_grid; return B200_ERR_INVALID; }
    // the reference decoder range-checks slice_y against num_h_slices (FFV1_Slice.cpp:127) and needs
    // num_h_slices < width, num_v_slices < height (FFV1_Frame.cpp:161-164)
    if (S->num_v > S->num_h || (uint32_t)S->num_h >= width || (uint32_t)S->num_v >= height) { *err = e_grid; return B200_ERR_INVALID; }

    QuantSets* Q = new QuantSets;
    make_quant_sets(S->bits, Q);
    std::memcpy(S->qtab, Q->q[S->context], sizeof S->qtab);
    S->nctx = Q->nctx[S->context];
    S->is5 = S->qtab[3][127] != 0;
    std::memcpy(S->one_state, kCustomOneState, 256);
    S->zero_state[0] = 0;
    for (int i = 1; i < 256; i++) S->zero_state[i] = (uint8_t)(256 - S->one_state[256 - i]);

    S->slices.clear();
    for (int sy = 0; sy < S->num_v; sy++)
        for (int sx = 0; sx < S->num_h; sx++) {
            SliceGeom g;
            g.x0 = (int32_t)((uint64_t)sx * width / S->num_h);
            g.y0 = (int32_t)((uint64_t)sy * height / S->num_v);
            g.w = (int32_t)((uint64_t)(sx + 1) * width / S->num_h) - g.x0;
            g.h = (int32_t)((uint64_t)(sy + 1) * height / S->num_v) - g.y0;
            S->slices.push_back(g);
        }

    // ---- ConfigurationRecord (order of fields = parameters::Parse, FFV1_Parameters.cpp:23-183)
    uint8_t def[256];
    default_one_state(def);
    {
        HostRangeEncoder rc(def);
        uint8_t st[32];
        std::memset(st, 128, sizeof st);
        rc.symbol(st, 3, false);                         // version
        rc.symbol(st, 4, false);                         // micro_version
        rc.symbol(st, 2, false);                         // coder_type: range coder, custom transitions
        for (int i = 1; i < 256; i++) rc.symbol(st, (int)kCustomOneState[i] - (int)def[i], true);
        rc.symbol(st, 1, false);                         // colorspace_type: JPEG2000-RCT
        rc.symbol(st, S->bits, false);                   // bits_per_raw_sample
        rc.bin(st[0], true);                             // chroma_planes
        rc.symbol(st, 0, false);                         // log2_h_chroma_subsample
        rc.symbol(st, 0, false);                         // log2_v_chroma_subsample
        rc.bin(st[0], false);                            // alpha_plane
        rc.symbol(st, S->num_h - 1, false);
        rc.symbol(st, S->num_v - 1, false);
        rc.symbol(st, 2, false);                         // quant_table_set_count
        for (int s = 0; s < 2; s++)
            for (int t = 0; t < 5; t++) put_qtab(rc, Q->q[s][t]);
        for (int s = 0; s < 2; s++) rc.bin(st[0], false);  // states_coded: all initial states are 128
        rc.symbol(st, S->ec, false);                     // ec
        rc.symbol(st, 1, false);                         // intra (-g 1)
        std::vector<uint8_t> rec = rc.finish();
        uint32_t crc = crc32_mpeg(rec.data(), rec.size());
        for (int k = 3; k >= 0; k--) rec.push_back((uint8_t)(crc >> (8 * k)));
        S->config_record = rec;
    }
    delete Q;

    // ---- bins ahead of the first sample of each slice: keyframe bin (first slice only, FFV1_Frame.cpp:148-156) and the
    // slice header symbols on one shared 32-state set (FFV1_Slice.cpp:113-177). States evolve deterministically, so the
    // (state, bit) pairs can be recorded here once and replayed by the GPU range coder.
    S->header_bins.clear();
    for (int sy = 0; sy < S->num_v; sy++)
        for (int sx = 0; sx < S->num_h; sx++) {
            HostRangeEncoder rc(S->one_state);
            if (sx == 0 && sy == 0) { uint8_t key = 128; rc.bin(key, true); }
            uint8_t st[32];
            std::memset(st, 128, sizeof st);
            rc.symbol(st, sx, false);
            rc.symbol(st, sy, false);
            rc.symbol(st, 0, false);                     // slice_width - 1 (grid units)
            rc.symbol(st, 0, false);                     // slice_height - 1
            rc.symbol(st, S->context, false);            // quant_table_set_index, plane-set 0 (Y)
            rc.symbol(st, S->context, false);            // plane-set 1 (Cb, Cr)
            rc.symbol(st, 3, false);                     // picture_structure: progressive
            rc.symbol(st, 0, false);                     // sar_num
            rc.symbol(st, 1, false);                     // sar_den
            S->header_bins.push_back(rc.recorded);
        }
    return 0;
}

}  // namespace b200
