/* oracle/flac_oracle.c — TEST INFRASTRUCTURE ONLY. Never linked into, loaded or called by the product.
 *
 * Plain-C restatement of the FLAC per-block ENCODE path of the B200 build: the path RAWcooked
 * hands to ffmpeg with `-c:a flac` (/root/reference/Source/CLI/Global.cpp:949-950). The bitstream syntax is the exact inverse
 * of the reference's vendored libFLAC 1.3.2 decoder:
 *   frame header        read_frame_header_   Source/Lib/ThirdParty/flac/src/libFLAC/stream_decoder.c:2159-2466
 *   subframe header     read_subframe_       stream_decoder.c:2468-2540
 *   fixed predictor     FLAC__fixed_restore_signal / FLAC__fixed_compute_residual   fixed.c:336-395
 *   partitioned Rice    read_residual_partitioned_rice_   stream_decoder.c:2745-2788, bitreader.c:744
 *   frame CRC-8/CRC-16  crc.c:366-376, checked at stream_decoder.c:2075-2127
 *   STREAMINFO          read_metadata_streaminfo_   stream_decoder.c:1565
 * Encoder decisions (deterministic): independent channels; per subframe CONSTANT if all samples are equal, else the smaller of
 *   - the fixed predictor order 0..4 with the smallest sum of |residual| over samples 4..n-1 (fixed.c:217-273 criterion),
 *   - an LPC predictor of order 1..8 with 15-bit coefficients (ffmpeg's level 5; analysis described at lpc_analyse below),
 * each with the Rice parameter per partition = floor(log2(mean)) and the partition order = exact minimum over 0..8; VERBATIM
 * when neither is smaller than the raw samples.
 * FFmpeg's own choices differ in detail: bitstream identity with FFmpeg is NOT pinned for FLAC (nothing in the reference pins it,
 * SURVEY.md §8c). What is pinned: the reference's libFLAC decodes every frame to the input PCM and verifies the STREAMINFO MD5
 * (tests/test_flac.py through oracle/_ref) — the property the reference's own tests check (test2.sh, check.sh) — and the output
 * is not larger than 1.02 x libavcodec's on the benchmark signal (tests/test_flac.py::test_oracle_size_next_to_ffmpeg_flac).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t* buf; size_t cap; uint64_t bitpos; int overflow; } bitw;

static void bw_put(bitw* w, uint32_t v, int n)          /* n <= 32, MSB first */
{
    for (int i = n - 1; i >= 0; i--) {
        size_t byte = (size_t)(w->bitpos >> 3);
        if (byte >= w->cap) { w->overflow = 1; w->bitpos++; continue; }
        if ((v >> i) & 1u) w->buf[byte] |= (uint8_t)(0x80 >> (w->bitpos & 7));
        w->bitpos++;
    }
}
static void bw_zeros(bitw* w, uint64_t n) { w->bitpos += n; if ((w->bitpos >> 3) > w->cap) w->overflow = 1; }

static uint8_t crc8(const uint8_t* d, size_t n)          /* poly 0x07, init 0 (crc.c) */
{
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) { c ^= d[i]; for (int k = 0; k < 8; k++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : c << 1); }
    return c;
}
static uint16_t crc16(const uint8_t* d, size_t n)        /* poly 0x8005, init 0 (crc.c) */
{
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) { c ^= (uint16_t)(d[i] << 8); for (int k = 0; k < 8; k++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : c << 1); }
    return c;
}

/* ffmpeg-like block size: the largest standard size not above 105 ms of audio */
int flaco_blocksize(int sample_rate)
{
    static const int sizes[] = {192, 256, 512, 576, 1024, 1152, 2048, 2304, 4096, 4608, 8192, 16384};
    int target = (int)((int64_t)sample_rate * 105 / 1000), best = 192;
    for (unsigned i = 0; i < sizeof sizes / sizeof sizes[0]; i++) if (sizes[i] <= target && sizes[i] > best) best = sizes[i];
    return best;
}

static int blocksize_code(int n, int* extra_bits)
{
    *extra_bits = 0;
    if (n == 192) return 1;
    for (int k = 0; k < 4; k++) if (n == 576 << k) return 2 + k;
    for (int k = 0; k < 8; k++) if (n == 256 << k) return 8 + k;
    if (n <= 256) { *extra_bits = 8; return 6; }
    *extra_bits = 16; return 7;
}
static int samplerate_code(int r, int* extra_bits, uint32_t* extra)
{
    static const int std_r[] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    *extra_bits = 0; *extra = 0;
    for (int i = 1; i < 12; i++) if (r == std_r[i]) return i;
    if (r % 1000 == 0 && r / 1000 < 256) { *extra_bits = 8; *extra = (uint32_t)(r / 1000); return 12; }
    if (r < 65536) { *extra_bits = 16; *extra = (uint32_t)r; return 13; }
    if (r % 10 == 0 && r / 10 < 65536) { *extra_bits = 16; *extra = (uint32_t)(r / 10); return 14; }
    return 0;
}
static int samplesize_code(int bps) { return bps == 8 ? 1 : bps == 12 ? 2 : bps == 16 ? 4 : bps == 20 ? 5 : bps == 24 ? 6 : 0; }

static int utf8_put(bitw* w, uint64_t v)
{
    if (v < 0x80) { bw_put(w, (uint32_t)v, 8); return 1; }
    int n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    bw_put(w, (uint32_t)((0xFF00u >> n) & 0xFF) | (uint32_t)(v >> (6 * (n - 1))), 8);
    for (int i = n - 2; i >= 0; i--) bw_put(w, 0x80 | (uint32_t)((v >> (6 * i)) & 0x3F), 8);
    return n;
}

/* ---- partitioned Rice: exact search of the partition order for the residual u[pred_order..n) (zig-zag folded) ---- */
typedef struct { uint64_t cost; int p, rice2; uint8_t k[256]; } rice_choice;

static void rice_search(const uint32_t* u, int n, int pred_order, rice_choice* best)
{
    int pmax = 0;
    while (pmax < 8 && (n % (2 << pmax)) == 0 && (n >> (pmax + 1)) > pred_order) pmax++;
    best->cost = UINT64_MAX; best->p = 0; best->rice2 = 0;
    for (int p = 0; p <= pmax; p++) {
        const int np = 1 << p, m = n >> p;
        uint64_t cost = 0; int anybig = 0; uint8_t ks[256];
        for (int j = 0; j < np; j++) {
            const int b = j ? j * m : pred_order, end = (j + 1) * m, cnt = end - b;
            uint64_t S = 0;
            for (int i = b; i < end; i++) S += u[i];
            int k = 0;
            while (k < 30 && ((uint64_t)cnt << (k + 1)) <= S) k++;
            uint64_t bits = (uint64_t)cnt * (uint64_t)(k + 1);
            for (int i = b; i < end; i++) bits += u[i] >> k;
            ks[j] = (uint8_t)k; cost += bits; if (k > 14) anybig = 1;
        }
        cost += (uint64_t)np * (anybig ? 5 : 4);
        if (cost < best->cost) { best->cost = cost; best->p = p; best->rice2 = anybig; memcpy(best->k, ks, (size_t)np); }
    }
}

static void rice_put(bitw* w, const uint32_t* u, int n, int pred_order, const rice_choice* c)
{
    bw_put(w, c->rice2 ? 1 : 0, 2);
    bw_put(w, (uint32_t)c->p, 4);
    const int np = 1 << c->p, m = n >> c->p;
    for (int j = 0; j < np; j++) {
        const int b = j ? j * m : pred_order, end = (j + 1) * m, k = c->k[j];
        bw_put(w, (uint32_t)k, c->rice2 ? 5 : 4);
        for (int i = b; i < end; i++) {
            bw_zeros(w, u[i] >> k);
            bw_put(w, 1, 1);
            if (k) bw_put(w, u[i] & ((1u << k) - 1), k);
        }
    }
}

/* ---- LPC analysis. What the decoder fixes is the predictor arithmetic
 * (FLAC__lpc_restore_signal / _wide, lpc.c:784-1100: data[i] = residual[i] + (int32)(sum(qlp[j] * data[i-j-1]) >> shift), 64-bit
 * sum whenever bps + precision + ilog2(order) > 32, stream_decoder.c:2710-2716) and the subframe syntax
 * (read_subframe_lpc_, stream_decoder.c:2627-2722). The analysis is the encoder's own business; this one is built to give the
 * same bits on a CPU and on a GPU: integer Welch window, exact 64-bit integer autocorrelation, Levinson-Durbin in IEEE double
 * with every operation a separate correctly rounded +, -, *, / (no fused multiply-add: compile with -ffp-contract=off),
 * coefficient quantisation as in FLAC__lpc_quantize_coefficients (lpc.c:166-266: 15-bit precision, error feedback), orders
 * 1..8 like ffmpeg's level 5, the order picked by an integer estimate of the Rice-coded size. */
#define LPC_MAX_ORDER 8
#define LPC_PRECISION 15

static int lpc_window15(int i, int n) { return (int)((4 * (int64_t)i * (int64_t)(n - 1 - i) * 32767) / ((int64_t)(n - 1) * (int64_t)(n - 1))); }

typedef struct { int order, shift; int32_t q[LPC_MAX_ORDER]; } lpc_params;

/* quantised predictors of orders 1..LPC_MAX_ORDER for x[0..n); valid[m-1] = 0 when order m is unusable. Returns 0 if the block
 * has no energy (nothing to analyse). */
static int lpc_analyse(const int32_t* x, int n, lpc_params out[LPC_MAX_ORDER], int valid[LPC_MAX_ORDER])
{
    int64_t ac[LPC_MAX_ORDER + 1];
    for (int l = 0; l <= LPC_MAX_ORDER; l++) ac[l] = 0;
    int32_t* xw = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) xw[i] = (int32_t)(((int64_t)x[i] * lpc_window15(i, n)) >> 15);
    for (int l = 0; l <= LPC_MAX_ORDER; l++)
        for (int i = l; i < n; i++) ac[l] += (int64_t)xw[i] * xw[i - l];
    free(xw);
    for (int m = 0; m < LPC_MAX_ORDER; m++) valid[m] = 0;
    if (ac[0] <= 0) return 0;
    double R[LPC_MAX_ORDER + 1], a[LPC_MAX_ORDER], prev[LPC_MAX_ORDER];
    for (int l = 0; l <= LPC_MAX_ORDER; l++) R[l] = (double)ac[l];
    double E = R[0];
    for (int m = 1; m <= LPC_MAX_ORDER; m++) {
        double acc = R[m];
        for (int j = 1; j < m; j++) { const double t = prev[j - 1] * R[m - j]; acc = acc - t; }
        if (!(E > 0.0)) break;
        const double k = acc / E;
        a[m - 1] = k;
        for (int j = 1; j < m; j++) { const double t = k * prev[m - 1 - j]; a[j - 1] = prev[j - 1] - t; }
        { const double t = k * k; const double u = 1.0 - t; E = E * u; }
        for (int j = 0; j < m; j++) prev[j] = a[j];
        /* quantise order m */
        double cmax = 0.0;
        for (int j = 0; j < m; j++) { const double d = a[j] < 0.0 ? -a[j] : a[j]; if (d > cmax) cmax = d; }
        if (!(cmax > 0.0) || !(cmax < 1.0e9)) continue;
        uint64_t bits; memcpy(&bits, &cmax, 8);
        const int e = (int)((bits >> 52) & 0x7FF) - 1022;             /* cmax = f * 2^e, f in [0.5, 1): frexp's exponent */
        int shift = LPC_PRECISION - 1 - e;
        if (shift > 15) shift = 15;
        if (shift < 0) continue;
        const int64_t qmax = (1 << (LPC_PRECISION - 1)) - 1, qmin = -(1 << (LPC_PRECISION - 1));
        const double scale = (double)(1 << shift);
        double err = 0.0;
        lpc_params* P = &out[m - 1];
        P->order = m; P->shift = shift;
        for (int j = 0; j < m; j++) {
            const double t = a[j] * scale;
            err = err + t;
            const double r = err >= 0.0 ? err + 0.5 : err - 0.5;
            int64_t q = (int64_t)r;                                   /* truncation: round half away from zero */
            if (q > qmax) q = qmax;
            if (q < qmin) q = qmin;
            err = err - (double)q;
            P->q[j] = (int32_t)q;
        }
        valid[m - 1] = 1;
    }
    return 1;
}

/* residual of predictor P over x[P->order..n) into u (zig-zag); returns 0 if a residual does not fit 31 bits. sum_u = sum of u. */
static int lpc_residual(const int32_t* x, int n, const lpc_params* P, uint32_t* u, uint64_t* sum_u)
{
    uint64_t S = 0;
    for (int i = P->order; i < n; i++) {
        int64_t sum = 0;
        for (int j = 0; j < P->order; j++) sum += (int64_t)P->q[j] * (int64_t)x[i - 1 - j];
        const int32_t pred = (int32_t)(sum >> P->shift);
        const int64_t e = (int64_t)x[i] - (int64_t)pred;
        if (e > 0x3FFFFFFF || e < -0x40000000) return 0;
        const int32_t e32 = (int32_t)e;
        const uint32_t z = ((uint32_t)e32 << 1) ^ (uint32_t)(e32 >> 31);
        if (u) u[i] = z;
        S += z;
    }
    *sum_u = S;
    return 1;
}

/* subframe of one channel: x[0..n) (already signed, bps bits) */
static void encode_subframe(bitw* w, const int32_t* x, int n, int bps, int use_lpc)
{
    const uint32_t vmask = bps == 32 ? 0xFFFFFFFFu : ((1u << bps) - 1);
    int constant = 1;
    for (int i = 1; i < n; i++) if (x[i] != x[0]) { constant = 0; break; }
    if (constant) { bw_put(w, 0x00, 8); bw_put(w, (uint32_t)x[0] & vmask, bps); return; }
    int order = -1;
    rice_choice fixed_c, lpc_c;
    uint64_t fixed_bits = UINT64_MAX, lpc_bits = UINT64_MAX;
    lpc_params lp; lp.order = 0; lp.shift = 0;
    uint32_t* u = NULL;
    if (n >= 16) {
        uint64_t sum[5] = {0, 0, 0, 0, 0};
        for (int i = 4; i < n; i++) {
            int64_t e0 = x[i], e1 = e0 - x[i - 1], e2 = e1 - ((int64_t)x[i - 1] - x[i - 2]);
            int64_t d2p = (int64_t)x[i - 1] - 2 * (int64_t)x[i - 2] + x[i - 3];
            int64_t e3 = e2 - d2p;
            int64_t d3p = (int64_t)x[i - 1] - 3 * (int64_t)x[i - 2] + 3 * (int64_t)x[i - 3] - x[i - 4];
            int64_t e4 = e3 - d3p;
            sum[0] += (uint64_t)(e0 < 0 ? -e0 : e0); sum[1] += (uint64_t)(e1 < 0 ? -e1 : e1); sum[2] += (uint64_t)(e2 < 0 ? -e2 : e2);
            sum[3] += (uint64_t)(e3 < 0 ? -e3 : e3); sum[4] += (uint64_t)(e4 < 0 ? -e4 : e4);
        }
        order = 0;
        for (int o = 1; o <= 4; o++) if (sum[o] < sum[order]) order = o;
        u = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
        /* ---- LPC candidate: the order whose estimated size is smallest, then its exact size */
        if (use_lpc && n > 2 * LPC_MAX_ORDER) {
            lpc_params cand[LPC_MAX_ORDER]; int valid[LPC_MAX_ORDER];
            if (lpc_analyse(x, n, cand, valid)) {
                uint64_t best_est = UINT64_MAX; int best_m = 0;
                for (int m = 1; m <= LPC_MAX_ORDER; m++) {
                    if (!valid[m - 1]) continue;
                    uint64_t S;
                    if (!lpc_residual(x, n, &cand[m - 1], NULL, &S)) continue;
                    const uint64_t cnt = (uint64_t)(n - m);
                    int k = 0;
                    while (k < 30 && (cnt << (k + 1)) <= S) k++;
                    const uint64_t est = cnt * (uint64_t)(k + 1) + (S >> k) + (uint64_t)m * (uint64_t)(bps + LPC_PRECISION);
                    if (est < best_est) { best_est = est; best_m = m; }
                }
                if (best_m) {
                    lp = cand[best_m - 1];
                    uint64_t S;
                    lpc_residual(x, n, &lp, u, &S);
                    rice_search(u, n, lp.order, &lpc_c);
                    lpc_bits = 8 + (uint64_t)lp.order * bps + 4 + 5 + (uint64_t)lp.order * LPC_PRECISION + 6 + lpc_c.cost;
                }
            }
        }
        /* ---- fixed candidate */
        for (int i = order; i < n; i++) {
            int64_t r;
            switch (order) {
            case 0: r = x[i]; break;
            case 1: r = (int64_t)x[i] - x[i - 1]; break;
            case 2: r = (int64_t)x[i] - 2 * (int64_t)x[i - 1] + x[i - 2]; break;
            case 3: r = (int64_t)x[i] - 3 * (int64_t)x[i - 1] + 3 * (int64_t)x[i - 2] - x[i - 3]; break;
            default: r = (int64_t)x[i] - 4 * (int64_t)x[i - 1] + 6 * (int64_t)x[i - 2] - 4 * (int64_t)x[i - 3] + x[i - 4]; break;
            }
            const int32_t e = (int32_t)r;
            u[i] = ((uint32_t)e << 1) ^ (uint32_t)(e >> 31);
        }
        rice_search(u, n, order, &fixed_c);
        fixed_bits = 8 + (uint64_t)order * bps + 6 + fixed_c.cost;
    }
    const uint64_t verbatim_bits = 8 + (uint64_t)n * bps;
    if (lpc_bits < fixed_bits && lpc_bits < verbatim_bits) {
        uint64_t S;
        lpc_residual(x, n, &lp, u, &S);
        bw_put(w, (uint32_t)((32 | (lp.order - 1)) << 1), 8);   /* LPC: 0 | 1ooooo | 0 (read_subframe_, stream_decoder.c:2521-2527) */
        for (int i = 0; i < lp.order; i++) bw_put(w, (uint32_t)x[i] & vmask, bps);
        bw_put(w, LPC_PRECISION - 1, 4);
        bw_put(w, (uint32_t)lp.shift, 5);
        for (int j = 0; j < lp.order; j++) bw_put(w, (uint32_t)lp.q[j] & ((1u << LPC_PRECISION) - 1), LPC_PRECISION);
        rice_put(w, u, n, lp.order, &lpc_c);
    } else if (order < 0 || fixed_bits >= verbatim_bits) {
        bw_put(w, 0x02, 8);                                   /* VERBATIM: 0 | 000001 | 0 */
        for (int i = 0; i < n; i++) bw_put(w, (uint32_t)x[i] & vmask, bps);
    } else {
        bw_put(w, (uint32_t)((8 | order) << 1), 8);           /* FIXED: 0 | 001ooo | 0 */
        for (int i = 0; i < order; i++) bw_put(w, (uint32_t)x[i] & vmask, bps);
        rice_put(w, u, n, order, &fixed_c);
    }
    free(u);
}

/* one frame: pcm = interleaved int32 [n][channels]; returns bytes written (0 on overflow) */
size_t flaco_encode_frame2(const int32_t* pcm, int n, int channels, int bps, int sample_rate, int use_lpc,
                           uint64_t frame_number, uint8_t* out, size_t cap);
size_t flaco_encode_frame(const int32_t* pcm, int n, int channels, int bps, int sample_rate, int blocksize_nominal,
                          uint64_t frame_number, uint8_t* out, size_t cap)
{
    (void)blocksize_nominal;
    return flaco_encode_frame2(pcm, n, channels, bps, sample_rate, 1, frame_number, out, cap);
}

/* use_lpc = 0: fixed predictors only (`compression_level` below 3 in ffmpeg's terms) */
size_t flaco_encode_frame2(const int32_t* pcm, int n, int channels, int bps, int sample_rate, int use_lpc,
                           uint64_t frame_number, uint8_t* out, size_t cap)
{
    memset(out, 0, cap);
    bitw w = {out, cap, 0, 0};
    int bs_extra, sr_extra; uint32_t sr_val;
    const int bsc = blocksize_code(n, &bs_extra), src = samplerate_code(sample_rate, &sr_extra, &sr_val);
    bw_put(&w, 0xFFF8, 16);                                   /* sync, reserved 0, fixed block size */
    bw_put(&w, (uint32_t)bsc, 4); bw_put(&w, (uint32_t)src, 4);
    bw_put(&w, (uint32_t)(channels - 1), 4); bw_put(&w, (uint32_t)samplesize_code(bps), 3); bw_put(&w, 0, 1);
    utf8_put(&w, frame_number);
    if (bs_extra) bw_put(&w, (uint32_t)(n - 1), bs_extra);
    if (sr_extra) bw_put(&w, sr_val, sr_extra);
    bw_put(&w, crc8(out, (size_t)(w.bitpos >> 3)), 8);
    int32_t* x = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int c = 0; c < channels; c++) {
        for (int i = 0; i < n; i++) x[i] = pcm[(size_t)i * channels + c];
        encode_subframe(&w, x, n, bps, use_lpc);
    }
    free(x);
    if (w.bitpos & 7) bw_zeros(&w, 8 - (w.bitpos & 7));
    if (w.overflow || (w.bitpos >> 3) + 2 > cap) return 0;
    const size_t nb = (size_t)(w.bitpos >> 3);
    const uint16_t c16 = crc16(out, nb);
    out[nb] = (uint8_t)(c16 >> 8); out[nb + 1] = (uint8_t)c16;
    return nb + 2;
}

/* "fLaC" + STREAMINFO metadata block = Matroska CodecPrivate of the A_FLAC track (fed to the decoder by
 * flac_wrapper::OutOfBand, Source/Lib/CoDec/Wrapper.cpp:138) */
size_t flaco_codec_private(int blocksize, uint32_t min_frame, uint32_t max_frame, int sample_rate, int channels, int bps,
                           uint64_t total_samples, uint8_t out[42])
{
    memset(out, 0, 42);
    memcpy(out, "fLaC", 4);
    out[4] = 0x80; out[5] = 0; out[6] = 0; out[7] = 34;
    uint8_t* s = out + 8;
    s[0] = (uint8_t)(blocksize >> 8); s[1] = (uint8_t)blocksize; s[2] = s[0]; s[3] = s[1];
    s[4] = (uint8_t)(min_frame >> 16); s[5] = (uint8_t)(min_frame >> 8); s[6] = (uint8_t)min_frame;
    s[7] = (uint8_t)(max_frame >> 16); s[8] = (uint8_t)(max_frame >> 8); s[9] = (uint8_t)max_frame;
    s[10] = (uint8_t)(sample_rate >> 12); s[11] = (uint8_t)(sample_rate >> 4);
    s[12] = (uint8_t)(((sample_rate & 15) << 4) | ((channels - 1) << 1) | (((bps - 1) >> 4) & 1));
    s[13] = (uint8_t)((((bps - 1) & 15) << 4) | (int)((total_samples >> 32) & 15));
    s[14] = (uint8_t)(total_samples >> 24); s[15] = (uint8_t)(total_samples >> 16); s[16] = (uint8_t)(total_samples >> 8); s[17] = (uint8_t)total_samples;
    /* MD5 of the unencoded audio: 0 = not computed (the decoder then skips the check) */
    return 42;
}
