/* oracle/ffv1_oracle.c — TEST INFRASTRUCTURE ONLY. Never linked into, loaded or called by the product.
 *
 * Plain-C, single-threaded, obviously-correct restatement of the FFV1 version-3 ENCODE path that
 * RAWcooked hands to ffmpeg (`-c:v ffv1 -coder 1 -context {0,1} -g 1 -level 3 -slicecrc 1 -slices N`,
 * /root/reference/Source/CLI/Global.cpp:938-989, command built at Source/CLI/Output.cpp:36-378).
 *
 * The arithmetic itself lives in a third-party dependency that is NOT in /root/reference:
 * FFmpeg libavcodec (ffv1enc.c / rangecoder.c), un-vendored and un-pinned by the reference; the copy in
 * this image is libavcodec 62.11.100 (FFmpeg 8.0).  This file restates that published algorithm (FFV1,
 * RFC 9043) as the exact inverse of the reference's in-tree decoder; each function cites the decoder
 * lines it inverts:
 *   range coder bin        inverse of rangecoder::b      Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:71-102
 *   symbol binarisation    inverse of rangecoder::u / s  FFV1_RangeCoder.cpp:105-171
 *   byte accounting        BytesUsed                     FFV1_RangeCoder.cpp:51-56
 *   configuration record   parameters::Parse             FFV1_Parameters.cpp:23-183, tables :222-253
 *   slice header / footer  slice::SliceHeader / Parse    FFV1_Slice.cpp:113-177, :247-253, :301-315
 *   plane/row order, borders  SliceContent_LineThenPlane FFV1_Slice.cpp:406-444
 *   predictor / context    predict, get_context_3/5      FFV1_Slice.cpp:21-93, Line :447-472
 *   frame assembly         ffv1_frame::Process           FFV1_Frame.cpp:148-197
 *   CRC                    ZenCRC32                      Source/Lib/Utils/CRC32/ZenCRC32.cpp:1097-1135
 *   pixel layouts + RCT    Transform                     Source/Lib/Transform/Transform.cpp:29-37, :70-420
 *
 * Pinning of the encoder half (tests/test_oracle.py): (1) every packet decodes through the UNMODIFIED reference decoder
 * (oracle/_ref/libref_ffv1dec.so) to the input bytes — the property all of the reference's own tests pin
 * (test1.sh, test2.sh, slices.sh: round trip through `rawcooked --check`); (2) config record and packets
 * are byte-identical to libavcodec 62.11.100's for the golden vectors in tests/golden/.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* ---- layouts: same numbering as include/b200enc.h (b200_layout) ---------------------------- */
enum {
    L_DPX_RGB_8 = 0, L_DPX_RGB_10_FA_LE = 1, L_DPX_RGB_10_FA_BE = 2, L_DPX_RGB_12_FA_LE = 3,
    L_DPX_RGB_12_PACKED_BE = 4, L_DPX_RGB_12_FA_BE = 5, L_DPX_RGB_16_LE = 6, L_DPX_RGB_16_BE = 7,
    L_TIFF_RGB_8 = 32, L_TIFF_RGB_16_LE = 33, L_TIFF_RGB_16_BE = 34
};

typedef struct {
    uint32_t width, height;
    int32_t layout;
    int32_t num_h, num_v;
    int32_t context;   /* 0 / 1 */
    int32_t ec;        /* slicecrc */
} ffv1o_cfg;

int ffv1o_layout_bits(int layout)
{
    switch (layout) {
    case L_DPX_RGB_8: case L_TIFF_RGB_8: return 8;
    case L_DPX_RGB_10_FA_LE: case L_DPX_RGB_10_FA_BE: return 10;
    case L_DPX_RGB_12_FA_LE: case L_DPX_RGB_12_FA_BE: case L_DPX_RGB_12_PACKED_BE: return 12;
    case L_DPX_RGB_16_LE: case L_DPX_RGB_16_BE: case L_TIFF_RGB_16_LE: case L_TIFF_RGB_16_BE: return 16;
    }
    return 0;
}

/* bytes per row of the payload as stored in the file (DPX rows are 32-bit aligned: DPX.cpp:478-482) */
size_t ffv1o_row_bytes(uint32_t w, int layout)
{
    switch (layout) {
    case L_DPX_RGB_8: return ((size_t)w * 3 + 3) & ~(size_t)3;
    case L_TIFF_RGB_8: return (size_t)w * 3;
    case L_DPX_RGB_10_FA_LE: case L_DPX_RGB_10_FA_BE: return (size_t)w * 4;
    case L_DPX_RGB_12_PACKED_BE: return (((size_t)w * 36 + 31) / 32) * 4;
    case L_DPX_RGB_16_LE: case L_DPX_RGB_16_BE: return ((size_t)w * 6 + 3) & ~(size_t)3;   /* not a Filled packing: DPX.cpp:478-482 */
    case L_DPX_RGB_12_FA_LE: case L_DPX_RGB_12_FA_BE: case L_TIFF_RGB_16_LE: case L_TIFF_RGB_16_BE: return (size_t)w * 6;
    }
    return 0;
}

size_t ffv1o_frame_bytes(uint32_t w, uint32_t h, int layout) { return ffv1o_row_bytes(w, layout) * h; }

/* read pixel (x,y) of the payload as the three stored components R,G,B (value bits only: padding bits
 * are the sidecar's business, DPX.cpp:500-608) */
static void load_rgb(const uint8_t* p, uint32_t w, int layout, uint32_t x, uint32_t y, int32_t* r, int32_t* g, int32_t* b)
{
    const uint8_t* row = p + ffv1o_row_bytes(w, layout) * y;
    switch (layout) {
    case L_DPX_RGB_8: case L_TIFF_RGB_8:
        *r = row[3 * x]; *g = row[3 * x + 1]; *b = row[3 * x + 2]; break;
    case L_DPX_RGB_10_FA_LE: case L_DPX_RGB_10_FA_BE: {
        const uint8_t* q = row + 4 * x;
        uint32_t v = layout == L_DPX_RGB_10_FA_BE ? ((uint32_t)q[0] << 24 | q[1] << 16 | q[2] << 8 | q[3])
                                                  : ((uint32_t)q[3] << 24 | q[2] << 16 | q[1] << 8 | q[0]);
        *r = (v >> 22) & 1023; *g = (v >> 12) & 1023; *b = (v >> 2) & 1023; break; }
    case L_DPX_RGB_12_FA_LE: case L_DPX_RGB_16_LE: case L_TIFF_RGB_16_LE: {
        const uint8_t* q = row + 6 * x;
        int sh = layout == L_DPX_RGB_12_FA_LE ? 4 : 0;
        *r = (q[0] | q[1] << 8) >> sh; *g = (q[2] | q[3] << 8) >> sh; *b = (q[4] | q[5] << 8) >> sh; break; }
    case L_DPX_RGB_12_FA_BE: case L_DPX_RGB_16_BE: case L_TIFF_RGB_16_BE: {
        const uint8_t* q = row + 6 * x;
        int sh = layout == L_DPX_RGB_12_FA_BE ? 4 : 0;
        *r = (q[0] << 8 | q[1]) >> sh; *g = (q[2] << 8 | q[3]) >> sh; *b = (q[4] << 8 | q[5]) >> sh; break; }
    case L_DPX_RGB_12_PACKED_BE: {
        /* component k of the row (k = 3x + c, order R,G,B) sits at bit 12k counted LSB-first inside
         * consecutive big-endian 32-bit words (Transform.cpp:217-316 read backwards) */
        int32_t c[3];
        for (int i = 0; i < 3; i++) {
            size_t bit = ((size_t)x * 3 + i) * 12;
            uint32_t v = 0;
            for (int k = 0; k < 12; k++) {
                size_t bb = bit + k;
                const uint8_t* q = row + (bb / 32) * 4;
                uint32_t word = (uint32_t)q[0] << 24 | q[1] << 16 | q[2] << 8 | q[3];
                v |= ((word >> (bb % 32)) & 1u) << k;
            }
            c[i] = (int32_t)v;
        }
        *r = c[0]; *g = c[1]; *b = c[2]; break; }
    default: *r = *g = *b = 0;
    }
}

/* ---- CRC-32, poly 0x04C11DB7, MSB first, init 0, no final xor (ZenCRC32.cpp:1097-1135 semantics) ---- */
static uint32_t crc_table[256];
static int crc_ready;
static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : c << 1;
        crc_table[i] = c;
    }
    crc_ready = 1;
}
uint32_t ffv1o_crc32(const uint8_t* d, size_t n)
{
    if (!crc_ready) crc_init();
    uint32_t c = 0;
    for (size_t i = 0; i < n; i++) c = (c << 8) ^ crc_table[(c >> 24) ^ d[i]];
    return c;
}

/* ---- state transition tables ---------------------------------------------------------------- */
/* default table of the range coder: FFmpeg's ff_build_rac_states(factor = 0.05 * 2^32, max_p = 256-8);
 * must equal the reference's default_state_transitions (FFV1_Frame.cpp:35-55) — tested. */
void ffv1o_default_transitions(uint8_t one_state[256])
{
    const int64_t one = (int64_t)1 << 32;
    const int factor = (int)(0.05 * (double)one);
    const int max_p = 256 - 8;
    int last_p8 = 0;
    int64_t p = one / 2;
    memset(one_state, 0, 256);
    for (int i = 0; i < 128; i++) {
        int p8 = (int)((256 * p + one / 2) >> 32);
        if (p8 <= last_p8) p8 = last_p8 + 1;
        if (last_p8 && last_p8 < 256 && p8 <= max_p) one_state[last_p8] = (uint8_t)p8;
        p += ((one - p) * factor + one / 2) >> 32;
        last_p8 = p8;
    }
    for (int i = 256 - max_p; i <= max_p; i++) {
        if (one_state[i]) continue;
        p = ((int64_t)i * one + 128) >> 8;
        p += ((one - p) * factor + one / 2) >> 32;
        int p8 = (int)((256 * p + one / 2) >> 32);
        if (p8 <= i) p8 = i + 1;
        if (p8 > max_p) p8 = max_p;
        one_state[i] = (uint8_t)p8;
    }
}

/* the custom table ffmpeg sends for `-coder 1` (coder_type 2 on the wire); observed through
 * parameters::Parse (FFV1_Parameters.cpp:41-55), SURVEY.md appendix A */
static const uint8_t custom_one_state[256] = {
      0, 10, 10, 10, 10, 16, 16, 16, 28, 16, 16, 29, 42, 49, 20, 49,
     59, 25, 26, 26, 27, 31, 33, 33, 33, 34, 34, 37, 67, 38, 39, 39,
     40, 40, 41, 79, 43, 44, 45, 45, 48, 48, 64, 50, 51, 52, 88, 52,
     53, 74, 55, 57, 58, 58, 74, 60,101, 61, 62, 84, 66, 66, 68, 69,
     87, 82, 71, 97, 73, 73, 82, 75,111, 77, 94, 78, 87, 81, 83, 97,
     85, 83, 94, 86, 99, 89, 90, 99,111, 92, 93,134, 95, 98,105, 98,
    105,110,102,108,102,118,103,106,106,113,109,112,114,112,116,125,
    115,116,117,117,126,119,125,121,121,123,145,124,126,131,127,129,
    165,130,132,138,133,135,145,136,137,139,146,141,143,142,144,148,
    147,155,151,149,151,150,152,157,153,154,156,168,158,162,161,160,
    172,163,169,164,166,184,167,170,177,174,171,173,182,176,180,178,
    175,189,179,181,186,183,192,185,200,187,191,188,190,197,193,196,
    197,194,195,196,198,202,199,201,210,203,207,204,205,206,208,214,
    209,211,221,212,213,215,224,216,217,218,219,220,222,228,223,225,
    226,224,227,229,240,230,231,232,233,234,235,236,238,239,237,242,
    241,243,242,244,245,246,247,248,249,250,251,252,252,253,254,255,
};
void ffv1o_custom_transitions(uint8_t one_state[256]) { memcpy(one_state, custom_one_state, 256); }

/* zero_state[i] = 256 - one_state[256 - i]  (AssignStateTransitions, FFV1_RangeCoder.cpp:36-42) */
static void make_zero_state(const uint8_t one[256], uint8_t zero[256])
{
    zero[0] = 0;
    for (int i = 1; i < 256; i++) zero[i] = (uint8_t)(256 - one[256 - i]);
}

/* ---- range ENcoder: inverse of rangecoder::b (FFV1_RangeCoder.cpp:71-102) ------------------- */
typedef void (*bin_hook)(void* user, uint8_t state, int bit);
typedef struct {
    uint8_t* buf; size_t cap, pos;
    uint32_t low, range;
    int outstanding_byte; size_t outstanding_count;
    uint8_t one[256], zero[256];
    int overflow;
    bin_hook hook; void* hook_user; uint64_t bins;
} renc;

static void renc_init(renc* c, uint8_t* buf, size_t cap, const uint8_t one[256])
{
    memset(c, 0, sizeof *c);
    c->buf = buf; c->cap = cap;
    c->low = 0; c->range = 0xFF00;           /* decoder: Mask = 0xFF then one refill (FFV1_RangeCoder.cpp:22-34,74-88) */
    c->outstanding_byte = -1;
    memcpy(c->one, one, 256);
    make_zero_state(one, c->zero);
}
static void renc_out(renc* c, int b)
{
    if (c->pos < c->cap) c->buf[c->pos] = (uint8_t)b; else c->overflow = 1;
    c->pos++;
}
static void renc_renorm(renc* c)
{
    while (c->range < 0x100) {
        if (c->outstanding_byte < 0) {
            c->outstanding_byte = (int)(c->low >> 8);
        } else if (c->low <= 0xFF00) {
            renc_out(c, c->outstanding_byte);
            for (; c->outstanding_count; c->outstanding_count--) renc_out(c, 0xFF);
            c->outstanding_byte = (int)(c->low >> 8);
        } else if (c->low >= 0x10000) {
            renc_out(c, c->outstanding_byte + 1);
            for (; c->outstanding_count; c->outstanding_count--) renc_out(c, 0x00);
            c->outstanding_byte = (int)((c->low >> 8) & 0xFF);
        } else {
            c->outstanding_count++;
        }
        c->low = (c->low & 0xFF) << 8;
        c->range <<= 8;
    }
}
static void renc_bin(renc* c, uint8_t* state, int bit)
{
    uint32_t r1 = (c->range * *state) >> 8;
    if (c->hook) c->hook(c->hook_user, *state, bit);
    c->bins++;
    if (!bit) { c->range -= r1; *state = c->zero[*state]; }
    else      { c->low += c->range - r1; c->range = r1; *state = c->one[*state]; }
    renc_renorm(c);
}
/* flush so that the decoder's BytesUsed() (FFV1_RangeCoder.cpp:51-56) lands exactly on the end */
static size_t renc_finish(renc* c)
{
    c->range = 0xFF; c->low += 0xFF; renc_renorm(c);
    c->range = 0xFF; renc_renorm(c);
    return c->pos;
}
/* inverse of rangecoder::u (:105-132) / rangecoder::s (:135-171) */
static void renc_symbol(renc* c, uint8_t* st, int32_t v, int is_signed)
{
    if (!v) { renc_bin(c, st + 0, 1); return; }
    uint32_t a = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    int e = 31 - __builtin_clz(a);
    renc_bin(c, st + 0, 0);
    for (int i = 0; i < e; i++) renc_bin(c, st + 1 + (i < 9 ? i : 9), 1);
    renc_bin(c, st + 1 + (e < 9 ? e : 9), 0);
    for (int i = e - 1; i >= 0; i--) renc_bin(c, st + 22 + (i < 9 ? i : 9), (a >> i) & 1);
    if (is_signed) renc_bin(c, st + 11 + (e < 10 ? e : 10), v < 0);
}

/* ---- quantisation tables (encoder's choice, carried in the record; FFV1_Parameters.cpp:222-253) ---- */
/* first 128 entries by threshold; mirrored like the decoder does */
static void fill_qt(int16_t t[256], const int* thr, int nthr, int scale)
{
    for (int i = 0; i < 128; i++) {
        int v = 0;
        for (int k = 0; k < nthr; k++) if (i >= thr[k]) v = k + 1;
        t[i] = (int16_t)(v * scale);
    }
    for (int i = 1; i < 128; i++) t[256 - i] = (int16_t)-t[i];
    t[128] = (int16_t)-t[127];
}
typedef struct { int16_t q[2][5][256]; int context_count[2]; } qtabs;
static void build_quant(int bits, qtabs* Q)
{
    static const int t11[] = {1, 2, 5, 12, 35};     /* 8-bit: 11 levels */
    static const int t5[]  = {1, 4};                /* 8-bit:  5 levels */
    static const int t9[]  = {5, 13, 27, 56};       /* >8-bit: 9 levels */
    static const int t5h[] = {11, 50};              /* >8-bit: 5 levels */
    memset(Q, 0, sizeof *Q);
    if (bits <= 8) {
        fill_qt(Q->q[0][0], t11, 5, 1); fill_qt(Q->q[0][1], t11, 5, 11); fill_qt(Q->q[0][2], t11, 5, 121);
        fill_qt(Q->q[1][0], t11, 5, 1); fill_qt(Q->q[1][1], t11, 5, 11);
        fill_qt(Q->q[1][2], t5, 2, 121); fill_qt(Q->q[1][3], t5, 2, 605); fill_qt(Q->q[1][4], t5, 2, 3025);
        Q->context_count[0] = (11 * 11 * 11 + 1) / 2; Q->context_count[1] = (11 * 11 * 5 * 5 * 5 + 1) / 2;
    } else {
        fill_qt(Q->q[0][0], t9, 4, 1); fill_qt(Q->q[0][1], t9, 4, 9); fill_qt(Q->q[0][2], t9, 4, 81);
        fill_qt(Q->q[1][0], t9, 4, 1); fill_qt(Q->q[1][1], t9, 4, 9);
        fill_qt(Q->q[1][2], t5h, 2, 81); fill_qt(Q->q[1][3], t5h, 2, 405); fill_qt(Q->q[1][4], t5h, 2, 2025);
        Q->context_count[0] = (9 * 9 * 9 + 1) / 2; Q->context_count[1] = (9 * 9 * 5 * 5 * 5 + 1) / 2;
    }
}
/* exported for the product-side tests: table `t` of set `s` */
void ffv1o_quant_table(int bits, int s, int t, int16_t out[256]) { qtabs Q; build_quant(bits, &Q); memcpy(out, Q.q[s][t], 512); }
int ffv1o_context_count(int bits, int s) { qtabs Q; build_quant(bits, &Q); return Q.context_count[s]; }

/* ---- `-slices N` -> grid (ffmpeg's search; valid set pinned by reference test/slices.sh:12) ---------- */
int ffv1o_slice_grid(uint32_t w, uint32_t h, int slices, int bits, int* num_h, int* num_v)
{
    int plane_count = 3;
    int v = (w > 352 || h > 288 || !slices) ? 2 : 1;
    if ((uint32_t)v > h) v = (int)h;
    for (; v <= 32; v++) {
        for (int hh = v; hh <= 2 * v; hh++) {
            int maxw = (int)((w + hh - 1) / hh), maxh = (int)((h + v - 1) / v);
            if ((uint32_t)hh > w || (uint32_t)v > h) continue;
            if ((int64_t)maxw * maxh * (bits + 1) * plane_count > (8 << 24)) continue;
            if (slices == hh * v && slices <= 1024) { *num_h = hh; *num_v = v; return 0; }
            if (maxw * maxh > 360 * 288) continue;
            if (!slices) { *num_h = hh; *num_v = v; return 0; }
        }
    }
    return -1;
}

/* ---- configuration record: what parameters::Parse reads (FFV1_Parameters.cpp:23-183) ---------- */
static void put_quant_table(renc* c, const int16_t* t)
{
    uint8_t st[32]; memset(st, 128, 32);
    int last = 0, i;
    for (i = 1; i < 128; i++)
        if (t[i] != t[i - 1]) { renc_symbol(c, st, i - last - 1, 0); last = i; }
    renc_symbol(c, st, i - last - 1, 0);
}
size_t ffv1o_config_record(const ffv1o_cfg* cfg, uint8_t* out, size_t cap)
{
    uint8_t def[256], st[32];
    uint8_t tmp[4096];
    int bits = ffv1o_layout_bits(cfg->layout);
    qtabs Q; build_quant(bits, &Q);
    ffv1o_default_transitions(def);
    renc c; renc_init(&c, tmp, sizeof tmp, def);
    memset(st, 128, 32);
    renc_symbol(&c, st, 3, 0);                 /* version                                  :27 */
    renc_symbol(&c, st, 4, 0);                 /* micro_version (>= 4 required)            :34-37 */
    renc_symbol(&c, st, 2, 0);                 /* coder_type 2 = range coder + custom table :38-55 */
    for (int i = 1; i < 256; i++) renc_symbol(&c, st, (int)custom_one_state[i] - (int)def[i], 1);
    renc_symbol(&c, st, 1, 0);                 /* colorspace_type 1 = JPEG2000-RCT         :57 */
    renc_symbol(&c, st, bits, 0);              /* bits_per_raw_sample                      :62 */
    renc_bin(&c, st, 1);                       /* chroma_planes                            :73 */
    renc_symbol(&c, st, 0, 0);                 /* log2_h_chroma_subsample                  :74 */
    renc_symbol(&c, st, 0, 0);                 /* log2_v_chroma_subsample                  :75 */
    renc_bin(&c, st, 0);                       /* alpha_plane                              :76 */
    renc_symbol(&c, st, cfg->num_h - 1, 0);    /* num_h_slices - 1                         :79 */
    renc_symbol(&c, st, cfg->num_v - 1, 0);    /* num_v_slices - 1                         :80 */
    renc_symbol(&c, st, 2, 0);                 /* quant_table_set_count                    :83 */
    for (int s = 0; s < 2; s++) for (int t = 0; t < 5; t++) put_quant_table(&c, Q.q[s][t]);
    for (int s = 0; s < 2; s++) renc_bin(&c, st, 0);   /* states_coded = 0 (all 128)        :101-103 */
    renc_symbol(&c, st, cfg->ec, 0);           /* ec                                       :136 */
    renc_symbol(&c, st, 1, 0);                 /* intra = 1 (-g 1)                         :143-149 */
    size_t n = renc_finish(&c);
    uint32_t crc = ffv1o_crc32(tmp, n);        /* CRC of record incl. parity == 0 (FFV1_Frame.cpp:114-117) */
    tmp[n] = crc >> 24; tmp[n + 1] = crc >> 16; tmp[n + 2] = crc >> 8; tmp[n + 3] = crc; n += 4;
    if (out) memcpy(out, tmp, n < cap ? n : cap);
    return n;
}

/* ---- one slice ---------------------------------------------------------------------------------- */
static int32_t median3(int32_t a, int32_t b, int32_t c)
{
    if (a > b) { int32_t t = a; a = b; b = t; }
    if (b > c) b = c;
    return a > b ? a : b;
}
static int32_t fold(int32_t d, int bits) { uint32_t m = 1u << (bits - 1); d &= (int32_t)((m << 1) - 1); return (int32_t)(((uint32_t)d ^ m) - m); }

typedef struct { uint64_t samples, bins; } ffv1o_slice_stats;

/* encode slice (sx, sy) into buf; returns slice byte count incl. footer.  `first` = slice at byte 0 of the
 * packet (carries the keyframe bin, FFV1_Frame.cpp:148-156, FFV1_Slice.cpp:217-221). */
static size_t encode_slice(const ffv1o_cfg* cfg, const uint8_t* payload, int sx, int sy, int first,
                           uint8_t* buf, size_t cap, bin_hook hook, void* hook_user, ffv1o_slice_stats* stats, int* overflow)
{
    const int bits = ffv1o_layout_bits(cfg->layout);
    const int sbits = bits <= 8 ? 9 : bits + 1;            /* bits_max for RCT, FFV1_Parameters.cpp:173-177 */
    const int32_t off = 1 << bits;
    const int swap_bg = bits > 8 && bits < 16;             /* "Exception indicated in specs": Transform.cpp:104,126,233,338,363 */
    const uint32_t W = cfg->width, H = cfg->height;
    const uint32_t x0 = (uint32_t)((uint64_t)sx * W / cfg->num_h), x1 = (uint32_t)((uint64_t)(sx + 1) * W / cfg->num_h);
    const uint32_t y0 = (uint32_t)((uint64_t)sy * H / cfg->num_v), y1 = (uint32_t)((uint64_t)(sy + 1) * H / cfg->num_v);
    const uint32_t w = x1 - x0, h = y1 - y0;               /* FFV1_Slice.cpp:153-156 */
    const int qidx = cfg->context ? 1 : 0;                 /* quant_table_set_index per plane-set */
    qtabs* Q = (qtabs*)malloc(sizeof(qtabs)); build_quant(bits, Q);
    const int16_t (*qt)[256] = Q->q[qidx];
    const int is5 = qt[3][127] != 0;                       /* FFV1_Slice.cpp:453 */
    const int nctx = Q->context_count[qidx];
    uint8_t def[256];
    ffv1o_default_transitions(def);

    renc c; renc_init(&c, buf, cap > 8 ? cap - 8 : 0, custom_one_state);
    c.hook = hook; c.hook_user = hook_user;
    if (first) { uint8_t ks = 128; renc_bin(&c, &ks, 1); }  /* keyframe = 1 */

    /* slice header, one shared 32-state set at 128 (FFV1_Slice.cpp:113-177) */
    uint8_t hs[32]; memset(hs, 128, 32);
    renc_symbol(&c, hs, sx, 0); renc_symbol(&c, hs, sy, 0);
    renc_symbol(&c, hs, 0, 0);  renc_symbol(&c, hs, 0, 0);          /* width-1, height-1 in grid units */
    renc_symbol(&c, hs, qidx, 0); renc_symbol(&c, hs, qidx, 0);      /* quant_table_set_index[0..1]    */
    renc_symbol(&c, hs, 3, 0);                                       /* picture_structure: progressive  */
    renc_symbol(&c, hs, 0, 0); renc_symbol(&c, hs, 1, 0);            /* sar 0/1 (unknown)               */

    /* context states: 2 plane-sets (Y | Cb+Cr), all 128 at every keyframe (FFV1_Coder_RangeCoder.cpp:25-48) */
    uint8_t* states = (uint8_t*)malloc((size_t)2 * nctx * 32);
    memset(states, 128, (size_t)2 * nctx * 32);

    /* 3 planes x 3 rows (cur, y-1, y-2), each with 2 guard samples left and 1 right, zero-initialised */
    const size_t stride = (size_t)w + 4;
    int32_t* rows = (int32_t*)calloc(3 * 3 * stride, sizeof(int32_t));
    #define ROW(p, k) (rows + ((size_t)(p) * 3 + (size_t)(k)) * stride + 2)
    int slot[3] = {0, 1, 2};                                 /* slot[0]=cur, [1]=prev, [2]=prev2 */
    uint64_t nsamples = 0;
    for (uint32_t y = 0; y < h; y++) {
        { int t = slot[2]; slot[2] = slot[1]; slot[1] = slot[0]; slot[0] = t; }
        for (uint32_t x = 0; x < w; x++) {
            int32_t r, g, b;
            load_rgb(payload, W, cfg->layout, x0 + x, y0 + y, &r, &g, &b);
            if (swap_bg) { int32_t t = g; g = b; b = t; }
            b -= g; r -= g; g += (b + r) >> 2; b += off; r += off;   /* forward of Transform.cpp:29-37 */
            ROW(0, slot[0])[x] = g; ROW(1, slot[0])[x] = b; ROW(2, slot[0])[x] = r;
        }
        for (int p = 0; p < 3; p++) {
            int32_t* cur = ROW(p, slot[0]); int32_t* prev = ROW(p, slot[1]); int32_t* prev2 = ROW(p, slot[2]);
            cur[-1] = prev[0];                               /* FFV1_Slice.cpp:430-431 */
            prev[w] = prev[w - 1];                           /* :432 */
            uint8_t* pst = states + (size_t)((p + 1) >> 1) * nctx * 32;
            for (uint32_t x = 0; x < w; x++) {
                int32_t L = cur[(int)x - 1], T = prev[x], LT = prev[(int)x - 1], RT = prev[x + 1];
                int32_t ctx = qt[0][(L - LT) & 255] + qt[1][(LT - T) & 255] + qt[2][(T - RT) & 255];
                if (is5) { int32_t LL = cur[(int)x - 2], TT = prev2[x]; ctx += qt[3][(LL - L) & 255] + qt[4][(TT - T) & 255]; }
                int32_t pred = median3(L, L + T - LT, T);
                int32_t d = cur[x] - pred;
                if (ctx < 0) { ctx = -ctx; d = -d; }
                d = fold(d, sbits);
                renc_symbol(&c, pst + (size_t)ctx * 32, d, 1);
                nsamples++;
            }
        }
    }
    #undef ROW
    { uint8_t es = 129; renc_bin(&c, &es, 0); }              /* terminator bin, FFV1_Slice.cpp:334-343 */
    size_t n = renc_finish(&c);
    if (stats) { stats->samples = nsamples; stats->bins = c.bins; }
    if (c.overflow || n + 8 > cap) { *overflow = 1; n = 0; }
    else {
        buf[n] = (uint8_t)(n >> 16); buf[n + 1] = (uint8_t)(n >> 8); buf[n + 2] = (uint8_t)n;   /* slice_size */
        n += 3;
        if (cfg->ec) {
            buf[n++] = 0;                                    /* error_status */
            uint32_t crc = ffv1o_crc32(buf, n);
            buf[n] = crc >> 24; buf[n + 1] = crc >> 16; buf[n + 2] = crc >> 8; buf[n + 3] = crc; n += 4;
        }
    }
    free(rows); free(states); free(Q);
    return n;
}

/* ---- one frame -> one packet ---------------------------------------------------------------------- */
/* returns packet size, 0 on overflow. slice_sizes (optional) gets num_h*num_v entries (bytes incl. footer). */
size_t ffv1o_encode_frame(const ffv1o_cfg* cfg, const uint8_t* payload, uint8_t* out, size_t cap,
                          uint32_t* slice_sizes, uint64_t* total_bins)
{
    size_t pos = 0; int overflow = 0; uint64_t bins = 0;
    for (int sy = 0; sy < cfg->num_v; sy++)
        for (int sx = 0; sx < cfg->num_h; sx++) {
            ffv1o_slice_stats st;
            size_t n = encode_slice(cfg, payload, sx, sy, pos == 0, out + pos, cap - pos, NULL, NULL, &st, &overflow);
            if (overflow) return 0;
            if (slice_sizes) slice_sizes[sy * cfg->num_h + sx] = (uint32_t)n;
            bins += st.bins;
            pos += n;
        }
    if (total_bins) *total_bins = bins;
    return pos;
}

/* ---- debugging aid for the CUDA path: the (state | bit<<8) sequence of one slice ------------------ */
typedef struct { uint16_t* out; size_t cap, n; } bin_dump;
static void dump_hook(void* u, uint8_t state, int bit)
{
    bin_dump* d = (bin_dump*)u;
    if (d->n < d->cap) d->out[d->n] = (uint16_t)(state | (bit << 8));
    d->n++;
}
size_t ffv1o_slice_bins(const ffv1o_cfg* cfg, const uint8_t* payload, int sx, int sy, uint16_t* bins, size_t cap)
{
    size_t bufcap = (size_t)cfg->width * cfg->height * 16 + 65536;
    uint8_t* buf = (uint8_t*)malloc(bufcap);
    bin_dump d = {bins, cap, 0};
    int overflow = 0;
    encode_slice(cfg, payload, sx, sy, sx == 0 && sy == 0, buf, bufcap, dump_hook, &d, NULL, &overflow);
    free(buf);
    return d.n;
}

/* =====================================================================================================
 * Decoder restatement (the `--check` side): plain C, one slice after the other, following
 *   record            parameters::Parse            Source/Lib/CoDec/FFV1/FFV1_Parameters.cpp:23-253
 *   packet -> slices  ffv1_frame::Process          Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:134-228
 *   slice             slice::Parse / SliceHeader   Source/Lib/CoDec/FFV1/FFV1_Slice.cpp:113-177, :210-318
 *   rows              SliceContent_LineThenPlane   FFV1_Slice.cpp:406-444, Line :447-472, predict / context :21-93
 *   range decoder     rangecoder::b / u / s        Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:25-171
 *   pixels            JPEG2000RCT + byte layouts   Source/Lib/Transform/Transform.cpp:29-37, :70-420
 * Pinned (tests/test_oracle.py): its output equals the UNMODIFIED reference decoder's (oracle/_ref) and the source payload
 * on every golden packet of libavcodec and on the encoder restatement's packets of every layout.
 * ===================================================================================================== */
typedef struct {
    const uint8_t* beg; size_t cur, end;      /* Buffer_Cur / Buffer_End as offsets */
    uint32_t low, range;                      /* Current / Mask */
    uint8_t one[256], zero[256];
    int underrun;
} rdec;

static void rdec_init(rdec* d, const uint8_t* buf, size_t size, const uint8_t one[256])
{
    d->beg = buf; d->cur = 0; d->end = size;
    d->low = size ? buf[0] : 0;               /* AssignBuffer, FFV1_RangeCoder.cpp:22-34 */
    d->range = 0xFF;
    d->cur = 1;
    d->underrun = 0;
    memcpy(d->one, one, 256);
    make_zero_state(d->one, d->zero);
}

static int rdec_bin(rdec* d, uint8_t* state)   /* rangecoder::b, FFV1_RangeCoder.cpp:71-102 */
{
    if (d->range < 0x100) {
        d->low <<= 8;
        if (d->cur > d->end) { d->underrun = 1; return 0; }
        if (d->cur < d->end) d->low |= d->beg[d->cur];
        d->range <<= 8;
        d->cur++;
    }
    uint32_t r1 = (d->range * *state) >> 8;
    d->range -= r1;
    if (d->low < d->range) { *state = d->zero[*state]; return 0; }
    d->low -= d->range;
    d->range = r1;
    *state = d->one[*state];
    return 1;
}

static uint32_t rdec_u(rdec* d, uint8_t* st)   /* rangecoder::u, :105-132 */
{
    if (rdec_bin(d, &st[0])) return 0;
    int e = 0;
    while (rdec_bin(d, &st[1 + (e < 9 ? e : 9)])) if (++e > 31) { d->underrun = 1; d->range = 0; return 0; }
    uint32_t a = 1;
    for (int i = e - 1; i >= 0; i--) a = (a << 1) | (uint32_t)rdec_bin(d, &st[22 + (i < 9 ? i : 9)]);
    return a;
}

static int32_t rdec_s(rdec* d, uint8_t* st)    /* rangecoder::s, :135-171 */
{
    if (rdec_bin(d, &st[0])) return 0;
    int e = 0;
    while (rdec_bin(d, &st[1 + (e < 9 ? e : 9)])) if (++e > 31) { d->underrun = 1; d->range = 0; return 0; }
    int32_t a = 1;
    for (int i = e - 1; i >= 0; i--) a = (a << 1) | rdec_bin(d, &st[22 + (i < 9 ? i : 9)]);
    return rdec_bin(d, &st[11 + (e < 10 ? e : 10)]) ? -a : a;
}

typedef struct {
    int bits, num_h, num_v, nsets, ec, intra;
    int nctx[8];
    int16_t q[8][5][256];
    uint8_t one[256];
} dec_params;

/* 0, or the reference's error text */
static const char* parse_record(const uint8_t* rec, size_t n, dec_params* P)
{
    if (n < 5) return "FFV1-HEADER-END:1";
    if (ffv1o_crc32(rec, n)) return "FFV1-HEADER-configuration_record_crc_parity:1";
    uint8_t def[256];
    ffv1o_default_transitions(def);
    rdec d;
    rdec_init(&d, rec, n - 4, def);
    uint8_t st[32];
    memset(st, 128, 32);
    if (rdec_u(&d, st) != 3) return "FFV1-HEADER-version";
    if (rdec_u(&d, st) < 4) return "FFV1-HEADER-micro_version-EXPERIMENTAL:1";
    uint32_t coder = rdec_u(&d, st);
    if (coder == 0 || coder > 2) return "FFV1-HEADER-coder_type:1";
    memcpy(P->one, def, 256);
    if (coder == 2)
        for (int i = 1; i < 256; i++) {
            int v = def[i] + rdec_s(&d, st);
            if (v < 0 || v > 255) return "FFV1-HEADER-state_transition_delta:1";
            P->one[i] = (uint8_t)v;
        }
    if (rdec_u(&d, st) != 1) return "colorspace_type";
    P->bits = (int)rdec_u(&d, st);
    if (!P->bits) P->bits = 8;
    if (!rdec_bin(&d, &st[0])) return "chroma_planes";
    if (rdec_u(&d, st) || rdec_u(&d, st)) return "chroma subsampling";
    if (rdec_bin(&d, &st[0])) return "alpha_plane";
    P->num_h = (int)rdec_u(&d, st) + 1;
    P->num_v = (int)rdec_u(&d, st) + 1;
    P->nsets = (int)rdec_u(&d, st);
    if (P->nsets < 1 || P->nsets > 8) return "FFV1-HEADER-quant_table_count:1";
    for (int s = 0; s < P->nsets; s++) {
        int scale = 1;
        for (int t = 0; t < 5; t++) {          /* parameters::QuantizationTable, FFV1_Parameters.cpp:222-253 */
            uint8_t qs[32];
            memset(qs, 128, 32);
            int v = 0;
            for (int k = 0; k < 128;) {
                uint32_t len_minus1 = rdec_u(&d, qs);
                if (k + len_minus1 >= 128 || d.underrun) return "FFV1-HEADER-QuantizationTable-len:1";
                for (uint32_t a = 0; a <= len_minus1; a++) P->q[s][t][k++] = (int16_t)(scale * v);
                v++;
            }
            for (int k = 1; k < 128; k++) P->q[s][t][256 - k] = (int16_t)-P->q[s][t][k];
            P->q[s][t][128] = (int16_t)-P->q[s][t][127];
            scale *= 2 * v - 1;
            if (scale > 32768) return "FFV1-HEADER-QuantizationTable-scale:1";
        }
        P->nctx[s] = (scale + 1) >> 1;
    }
    for (int s = 0; s < P->nsets; s++) if (rdec_bin(&d, &st[0])) return "states_coded";
    P->ec = (int)rdec_u(&d, st);
    if (P->ec > 1) return "FFV1-HEADER-ec:1";
    P->intra = (int)rdec_u(&d, st);
    return d.underrun ? "FFV1-HEADER-END:1" : 0;
}

/* inverse of load_rgb: writes pixel (x, y) into the payload (Transform.cpp:70-420) */
static void store_rgb(uint8_t* p, uint32_t w, int layout, uint32_t x, uint32_t y, uint32_t r, uint32_t g, uint32_t b)
{
    uint8_t* row = p + ffv1o_row_bytes(w, layout) * y;
    switch (layout) {
    case L_DPX_RGB_8: case L_TIFF_RGB_8:
        row[3 * x] = (uint8_t)r; row[3 * x + 1] = (uint8_t)g; row[3 * x + 2] = (uint8_t)b; break;
    case L_DPX_RGB_10_FA_LE: case L_DPX_RGB_10_FA_BE: {
        uint32_t v = r << 22 | g << 12 | b << 2;
        uint8_t* q = row + 4 * x;
        if (layout == L_DPX_RGB_10_FA_BE) { q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v; }
        else { q[3] = (uint8_t)(v >> 24); q[2] = (uint8_t)(v >> 16); q[1] = (uint8_t)(v >> 8); q[0] = (uint8_t)v; }
        break; }
    case L_DPX_RGB_12_FA_LE: case L_DPX_RGB_16_LE: case L_TIFF_RGB_16_LE: case L_DPX_RGB_12_FA_BE: case L_DPX_RGB_16_BE: case L_TIFF_RGB_16_BE: {
        int sh = (layout == L_DPX_RGB_12_FA_LE || layout == L_DPX_RGB_12_FA_BE) ? 4 : 0;
        int be = layout == L_DPX_RGB_12_FA_BE || layout == L_DPX_RGB_16_BE || layout == L_TIFF_RGB_16_BE;
        uint32_t c[3] = { r << sh, g << sh, b << sh };
        uint8_t* q = row + 6 * x;
        for (int i = 0; i < 3; i++) {
            if (be) { q[2 * i] = (uint8_t)(c[i] >> 8); q[2 * i + 1] = (uint8_t)c[i]; }
            else { q[2 * i + 1] = (uint8_t)(c[i] >> 8); q[2 * i] = (uint8_t)c[i]; }
        }
        break; }
    case L_DPX_RGB_12_PACKED_BE: {
        uint32_t c[3] = { r, g, b };
        for (int i = 0; i < 3; i++) {
            size_t bit = ((size_t)x * 3 + i) * 12;
            for (int k = 0; k < 12; k++) {
                size_t bb = bit + k;
                uint8_t* q = row + (bb / 32) * 4 + (3 - (bb % 32) / 8);   /* bit bb % 32 of a big-endian word */
                if ((c[i] >> k) & 1u) *q |= (uint8_t)(1u << (bb % 8)); else *q &= (uint8_t)~(1u << (bb % 8));
            }
        }
        break; }
    }
}

/* flags as in include/b200dec.h: 1 tail, 2 crc, 4 header, 8 underrun, 16 junk, 32 error_status. Returns 0 when the call worked
 * (the record parsed, the output buffer was large enough), else -1; *flags tells about the stream. Row padding is written as zeros. */
int ffv1o_decode_frame(const uint8_t* rec, size_t rec_len, uint32_t w, uint32_t h, int layout,
                       const uint8_t* pkt, size_t pkt_len, uint8_t* out, size_t out_cap, uint32_t* flags)
{
    dec_params* P = (dec_params*)malloc(sizeof(dec_params));
    *flags = 0;
    if (!P || parse_record(rec, rec_len, P) || P->bits != ffv1o_layout_bits(layout) || ffv1o_frame_bytes(w, h, layout) > out_cap) { free(P); return -1; }
    memset(out, 0, ffv1o_frame_bytes(w, h, layout));
    const size_t tail = P->ec ? 8 : 3;
    const int bits_max = P->bits + 1, off = 1 << P->bits, mask = (1 << bits_max) - 1;
    const int swap_bg = P->bits > 8 && P->bits < 16;          /* Transform.cpp:104,126,233,338,363 */
    int maxctx = 0;
    for (int s = 0; s < P->nsets; s++) if (P->nctx[s] > maxctx) maxctx = P->nctx[s];
    uint8_t* states = (uint8_t*)malloc((size_t)2 * maxctx * 32);
    size_t pos = pkt_len;
    int nslices = 0;
    while (pos) {                                             /* FFV1_Frame.cpp:166-197: slices from the end of the packet */
        if (pos < tail || nslices >= P->num_h * P->num_v) { *flags |= 1; break; }
        size_t size = ((size_t)pkt[pos - tail] << 16 | (size_t)pkt[pos - tail + 1] << 8 | pkt[pos - tail + 2]) + tail;
        if (size > pos) { *flags |= 1; break; }
        pos -= size;
        nslices++;
        const uint8_t* sp = pkt + pos;
        if (P->ec && ffv1o_crc32(sp, size)) *flags |= 2;      /* FFV1_Slice.cpp:247-249; decoding goes on */
        rdec d;
        rdec_init(&d, sp, size - tail, P->one);
        if (pos == 0) {                                       /* first slice: the keyframe bin, FFV1_Slice.cpp:221-225 */
            uint8_t key = 128;
            if (!rdec_bin(&d, &key)) { *flags |= 4; continue; }
        }
        uint8_t hs[32];
        memset(hs, 128, 32);
        uint32_t sx = rdec_u(&d, hs), sy = rdec_u(&d, hs), sw1 = rdec_u(&d, hs), sh1 = rdec_u(&d, hs);
        uint32_t qi[2] = { rdec_u(&d, hs), 0 };
        qi[1] = rdec_u(&d, hs);
        rdec_u(&d, hs); rdec_u(&d, hs); rdec_u(&d, hs);       /* picture_structure, sar_num, sar_den */
        uint32_t x2 = sx + sw1 + 1, y2 = sy + sh1 + 1;
        if (sx >= (uint32_t)P->num_h || sy >= (uint32_t)P->num_h || x2 > (uint32_t)P->num_h || y2 > (uint32_t)P->num_v ||
            qi[0] >= (uint32_t)P->nsets || qi[1] >= (uint32_t)P->nsets || d.underrun) { *flags |= 4; continue; }
        const uint32_t gx = (uint32_t)((uint64_t)sx * w / P->num_h), gy = (uint32_t)((uint64_t)sy * h / P->num_v);
        const uint32_t gw = (uint32_t)((uint64_t)x2 * w / P->num_h) - gx, gh = (uint32_t)((uint64_t)y2 * h / P->num_v) - gy;
        memset(states, 128, (size_t)2 * maxctx * 32);         /* keyframe: every state back to 128, FFV1_Coder_RangeCoder.cpp:25-48 */
        /* two rows per plane with 2 samples of margin on the left and 1 on the right (FFV1_Slice.cpp:406-444) */
        int32_t* buf = (int32_t*)calloc((size_t)3 * 2 * (gw + 3), sizeof(int32_t));
        int32_t* sample[3][2];
        for (int c = 0; c < 3; c++) { sample[c][0] = buf + (size_t)2 * c * (gw + 3) + 2; sample[c][1] = sample[c][0] + gw + 3; }
        for (uint32_t y = 0; y < gh; y++) {
            for (int c = 0; c < 3; c++) {
                int32_t* t = sample[c][0]; sample[c][0] = sample[c][1]; sample[c][1] = t;      /* swap */
                sample[c][1][-1] = sample[c][0][0];
                sample[c][0][gw] = sample[c][0][gw - 1];
                const int ps = (c + 1) >> 1;
                const int16_t (*Q)[256] = P->q[qi[ps]];
                const int is5 = Q[3][127] != 0;
                uint8_t* S = states + (size_t)ps * maxctx * 32;
                int32_t *s0 = sample[c][0], *s1 = sample[c][1];   /* previous row, current row (Line, :447-472) */
                for (uint32_t x = 0; x < gw; x++) {
                    const int LT = s0[(int)x - 1], T = s0[x], RT = s0[x + 1], L = s1[(int)x - 1];
                    int ctx = Q[0][(L - LT) & 255] + Q[1][(LT - T) & 255] + Q[2][(T - RT) & 255];
                    if (is5) ctx += Q[3][(s1[(int)x - 2] - L) & 255] + Q[4][(s1[x] - T) & 255];   /* LL, TT (row y-2 still in place) */
                    int pred = median3(L, L + T - LT, T);
                    int32_t v = ctx >= 0 ? pred + rdec_s(&d, S + (size_t)ctx * 32) : pred - rdec_s(&d, S + (size_t)(-ctx) * 32);
                    s1[x] = v & mask;
                }
            }
            for (uint32_t x = 0; x < gw; x++) {               /* JPEG2000RCT, Transform.cpp:29-37 */
                int32_t g = sample[0][1][x], b = sample[1][1][x] - off, r = sample[2][1][x] - off;
                g -= (b + r) >> 2; b += g; r += g;
                if (swap_bg) { int32_t t = g; g = b; b = t; }
                store_rgb(out, w, layout, gx + x, gy + y, (uint32_t)r & (uint32_t)(off - 1), (uint32_t)g & (uint32_t)(off - 1), (uint32_t)b & (uint32_t)(off - 1));
            }
        }
        free(buf);
        uint8_t endst = 129;                                  /* terminator, FFV1_Slice.cpp:335-341 */
        rdec_bin(&d, &endst);
        const size_t adj = d.range < 0x100 ? 0 : 1;
        if (d.underrun || d.cur - adj > d.end) *flags |= 8;
        const size_t used = d.cur > d.end ? d.end : d.cur - adj;
        if (used < d.end) *flags |= 16;
        if (P->ec && sp[size - 5]) *flags |= 32;
    }
    if (nslices != P->num_h * P->num_v) *flags |= 1;
    free(states);
    free(P);
    return 0;
}
