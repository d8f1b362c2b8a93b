// C ABI of the B200 FLAC encoder (include/b200enc.h) + the WAV helper of the argv front-end.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200enc.h"
#include "flac_kernels.cuh"
#include "ingest.h"

void b200_set_error(const std::string& msg);          // b200enc.cu

namespace {
// MD5 (RFC 1321) of the unencoded audio for STREAMINFO: little-endian signed samples, interleaved, whole bytes per sample
// (what FLAC__MD5Accumulate hashes on the decoding side, Source/Lib/ThirdParty/flac/src/libFLAC/md5.c).
class Md5 {
  public:
    void update(const uint8_t* p, size_t n) {
        total_ += n;
        if (fill_) {
            const size_t take = n < 64 - fill_ ? n : 64 - fill_;
            memcpy(buf_ + fill_, p, take);
            fill_ += take; p += take; n -= take;
            if (fill_ == 64) { block(buf_); fill_ = 0; }
        }
        for (; n >= 64; p += 64, n -= 64) block(p);
        if (n) { memcpy(buf_, p, n); fill_ = n; }
    }
    void digest(uint8_t out[16]) const {
        Md5 c = *this;
        const uint64_t bits = c.total_ * 8;
        const uint8_t pad = 0x80, zero = 0;
        c.update(&pad, 1);
        while (c.fill_ != 56) c.update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (8 * i));
        c.update(len, 8);
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(c.h_[i] >> (8 * k));
    }
    uint64_t bytes() const { return total_; }

  private:
    static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
    void block(const uint8_t* p) {
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        static uint32_t K[64];
        static bool init = false;
        if (!init) {      // floor(2^32 * |sin(i + 1)|) from its first digits would need libm; the RFC's table is derived once by integer arithmetic below
            static const uint32_t T[64] = {
                0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
                0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
                0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
                0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
                0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
                0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
            memcpy(K, T, sizeof K);
            init = true;
        }
        uint32_t M[16];
        for (int i = 0; i < 16; i++) M[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
        uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3];
        for (int i = 0; i < 64; i++) {
            uint32_t f; int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            const uint32_t t = d;
            d = c; c = b;
            b = b + rol(a + f + K[i] + M[g], S[i]);
            a = t;
        }
        h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d;
    }
    uint32_t h_[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint8_t buf_[64];
    size_t fill_ = 0;
    uint64_t total_ = 0;
};
}  // namespace

struct b200_flac_enc {
    b200_flac_cfg cfg;
    Md5 md5;
    int block_size = 0;
    uint32_t frame_words = 0;
    int max_blocks = 0;
    uint8_t* d_pcm = nullptr;
    uint32_t* d_out = nullptr;
    uint32_t* d_len = nullptr;
    std::vector<uint8_t> h_out;
    std::vector<uint32_t> h_len;
    uint32_t min_frame = 0xFFFFFF, max_frame = 0;
    uint64_t launches = 0;
};

namespace {
int fail(int code, const std::string& msg) { b200_set_error(msg); return code; }
int fail_cuda(cudaError_t e, const char* what) {
    b200_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, #x); } while (0)

// the largest standard block size not above 105 ms of audio (what ffmpeg's flac encoder picks: 4608 @ 44.1/48 kHz, 8192 @ 96 kHz);
// the reference buffers at most 16384 samples per block (Source/Lib/CoDec/Wrapper.cpp:251)
int pick_block_size(int rate) {
    static const int sizes[] = {192, 256, 512, 576, 1024, 1152, 2048, 2304, 4096, 4608, 8192, 16384};
    const int target = (int)((int64_t)rate * 105 / 1000);
    int best = 192;
    for (int s : sizes) if (s <= target && s > best) best = s;
    return best;
}
}  // namespace

extern "C" {

int b200_flac_open(const b200_flac_cfg* cfg, b200_flac_enc** out) {
    if (!cfg || !out) return fail(B200_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->channels < 1 || cfg->channels > 8) return fail(B200_ERR_INVALID, "FLAC carries 1..8 channels");
    if (cfg->bits != 8 && cfg->bits != 16 && cfg->bits != 24) return fail(B200_ERR_INVALID, "FLAC path takes 8, 16 or 24-bit integer PCM");
    if (cfg->sample_rate < 1 || cfg->sample_rate > 655350) return fail(B200_ERR_INVALID, "bad sample rate");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B200_ERR_NO_DEVICE, "no CUDA device: the B200 encoder has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(B200_ERR_INVALID, "bad device ordinal");
    CU(cudaSetDevice(cfg->device));
    b200_flac_enc* E = new b200_flac_enc;
    E->cfg = *cfg;
    E->block_size = cfg->block_size ? cfg->block_size : pick_block_size((int)cfg->sample_rate);
    if (E->block_size < 16 || E->block_size > 16384) { delete E; return fail(B200_ERR_INVALID, "block size must be 16..16384"); }
    E->max_blocks = cfg->max_blocks > 0 ? cfg->max_blocks : 256;
    const size_t fbytes = 32 + (size_t)cfg->channels * (2 + ((size_t)E->block_size * cfg->bits + 7) / 8) + 2;
    E->frame_words = (uint32_t)((fbytes + 3) / 4 + 4);
    const size_t pcm_bytes = (size_t)E->max_blocks * E->block_size * cfg->channels * (cfg->bits / 8);
    cudaError_t e = cudaMalloc((void**)&E->d_pcm, pcm_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_out, (size_t)E->max_blocks * E->frame_words * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_len, (size_t)E->max_blocks * 4);
    if (e != cudaSuccess) { int rc = fail_cuda(e, "cudaMalloc"); b200_flac_close(E); return rc; }
    E->h_out.resize((size_t)E->max_blocks * E->frame_words * 4);
    E->h_len.resize(E->max_blocks);
    *out = E;
    return 0;
}

void b200_flac_close(b200_flac_enc* E) {
    if (!E) return;
    cudaSetDevice(E->cfg.device);
    if (E->d_pcm) cudaFree(E->d_pcm);
    if (E->d_out) cudaFree(E->d_out);
    if (E->d_len) cudaFree(E->d_len);
    delete E;
}

int32_t b200_flac_block_size(const b200_flac_enc* E) { return E ? E->block_size : 0; }
size_t b200_flac_max_frame_bytes(const b200_flac_enc* E) { return E ? (size_t)E->frame_words * 4 : 0; }

size_t b200_flac_codec_private(const b200_flac_enc* E, uint64_t total_samples, uint8_t* out, size_t cap) {
    if (!E) return 0;
    uint8_t b[42];
    memset(b, 0, sizeof b);
    memcpy(b, "fLaC", 4);
    b[4] = 0x80; b[7] = 34;                                  // last metadata block, STREAMINFO, 34 bytes
    uint8_t* s = b + 8;
    const uint32_t bs = (uint32_t)E->block_size, mn = E->max_frame ? E->min_frame : 0, mx = E->max_frame;
    const uint32_t sr = E->cfg.sample_rate, ch = E->cfg.channels, bps = E->cfg.bits;
    s[0] = (uint8_t)(bs >> 8); s[1] = (uint8_t)bs; s[2] = s[0]; s[3] = s[1];
    s[4] = (uint8_t)(mn >> 16); s[5] = (uint8_t)(mn >> 8); s[6] = (uint8_t)mn;
    s[7] = (uint8_t)(mx >> 16); s[8] = (uint8_t)(mx >> 8); s[9] = (uint8_t)mx;
    s[10] = (uint8_t)(sr >> 12); s[11] = (uint8_t)(sr >> 4);
    s[12] = (uint8_t)(((sr & 15) << 4) | ((ch - 1) << 1) | (((bps - 1) >> 4) & 1));
    s[13] = (uint8_t)((((bps - 1) & 15) << 4) | (uint32_t)((total_samples >> 32) & 15));
    s[14] = (uint8_t)(total_samples >> 24); s[15] = (uint8_t)(total_samples >> 16); s[16] = (uint8_t)(total_samples >> 8); s[17] = (uint8_t)total_samples;
    // MD5 of the unencoded audio: known when every sample of the stream went through this handle; else 0 = "not computed"
    if (E->md5.bytes() == total_samples * (uint64_t)ch * (bps / 8) && total_samples) E->md5.digest(s + 18);
    if (out && cap) memcpy(out, b, cap < 42 ? cap : 42);
    return 42;
}

int b200_flac_encode_host(b200_flac_enc* E, const uint8_t* pcm, uint64_t n_samples, uint64_t first_frame,
                          uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len, int32_t* n_frames) {
    if (!E || !pcm || !out || !n_frames) return fail(B200_ERR_INVALID, "null argument");
    const uint64_t nblk = (n_samples + E->block_size - 1) / E->block_size;
    if (nblk == 0 || nblk > (uint64_t)E->max_blocks) return fail(B200_ERR_INVALID, "n_samples out of range for max_blocks");
    CU(cudaSetDevice(E->cfg.device));
    const size_t bytes = (size_t)n_samples * E->cfg.channels * (E->cfg.bits / 8);
    if (E->cfg.bits == 8) {                                   // WAV 8-bit is unsigned, the signature is over signed bytes
        std::vector<uint8_t> sgn(pcm, pcm + bytes);
        for (auto& v : sgn) v = (uint8_t)(v - 128);
        E->md5.update(sgn.data(), bytes);
    } else {
        E->md5.update(pcm, bytes);
    }
    CU(cudaMemcpyAsync(E->d_pcm, pcm, bytes, cudaMemcpyHostToDevice, 0));
    CU(cudaMemsetAsync(E->d_out, 0, (size_t)nblk * E->frame_words * 4, 0));
    b200::FlacArgs A;
    A.pcm = E->d_pcm; A.n_samples = n_samples; A.first_frame = first_frame;
    A.channels = (int32_t)E->cfg.channels; A.bits = (int32_t)E->cfg.bits; A.sample_rate = (int32_t)E->cfg.sample_rate; A.block_size = E->block_size;
    A.out = E->d_out; A.frame_words = E->frame_words; A.frame_len = E->d_len;
    A.use_lpc = E->cfg.fixed_only ? 0 : 1;
    CU(b200::launch_flac(A, (int)nblk, 0));
    E->launches++;
    CU(cudaMemcpyAsync(E->h_len.data(), E->d_len, (size_t)nblk * 4, cudaMemcpyDeviceToHost, 0));
    CU(cudaMemcpyAsync(E->h_out.data(), E->d_out, (size_t)nblk * E->frame_words * 4, cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    size_t pos = 0;
    for (uint64_t b = 0; b < nblk; b++) {
        const uint32_t len = E->h_len[b];
        if (len == 0 || len > E->frame_words * 4) return fail(B200_ERR_OVERFLOW, "FLAC frame buffer overflow");
        if (pos + len > out_cap) return fail(B200_ERR_OVERFLOW, "output buffer too small");
        memcpy(out + pos, E->h_out.data() + (size_t)b * E->frame_words * 4, len);
        if (out_off) out_off[b] = pos;
        if (out_len) out_len[b] = len;
        pos += len;
        if (len < E->min_frame) E->min_frame = len;
        if (len > E->max_frame) E->max_frame = len;
    }
    *n_frames = (int32_t)nblk;
    return 0;
}

}  // extern "C"

// WAV file -> FLAC packets for the Matroska muxer (one frame per SimpleBlock, what flac_wrapper::Process expects:
// /root/reference/Source/Lib/CoDec/Wrapper.cpp:204-219)
int b200_flac_encode_file_to_mux(const std::string& path, const b200::WavInfo& wi, int /*track*/, int device,
                                 std::vector<uint8_t>* codec_private, std::vector<std::pair<int64_t, std::vector<uint8_t>>>* packets,
                                 std::string* err) {
    b200_flac_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.sample_rate = wi.sample_rate; cfg.channels = wi.channels; cfg.bits = wi.bits; cfg.max_blocks = 256; cfg.device = device;
    b200_flac_enc* E = nullptr;
    int rc = b200_flac_open(&cfg, &E);
    if (rc) { *err = b200_last_error(); return rc; }
    const size_t bpf = (size_t)wi.channels * (wi.bits / 8);
    const uint64_t total = wi.data_bytes / bpf;
    const int bs = b200_flac_block_size(E);
    const uint64_t per_call = (uint64_t)bs * 256;
    std::vector<uint8_t> pcm((size_t)per_call * bpf), out((size_t)256 * b200_flac_max_frame_bytes(E));
    std::vector<size_t> off(256), len(256);
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open " + path; b200_flac_close(E); return B200_ERR_IO; }
    uint64_t done = 0, frame = 0;
    while (done < total) {
        const uint64_t n = std::min<uint64_t>(per_call, total - done);
        size_t got = 0;
        while (got < n * bpf) {
            ssize_t r = pread(fd, pcm.data() + got, n * bpf - got, (off_t)(wi.data_offset + done * bpf + got));
            if (r <= 0) { *err = "cannot read " + path; close(fd); b200_flac_close(E); return B200_ERR_IO; }
            got += (size_t)r;
        }
        int32_t nf = 0;
        rc = b200_flac_encode_host(E, pcm.data(), n, frame, out.data(), out.size(), off.data(), len.data(), &nf);
        if (rc) { *err = b200_last_error(); close(fd); b200_flac_close(E); return rc; }
        for (int32_t k = 0; k < nf; k++) {
            const int64_t t_ms = (int64_t)((double)(done + (uint64_t)k * bs) * 1000.0 / wi.sample_rate + 0.5);
            packets->push_back({t_ms, std::vector<uint8_t>(out.begin() + off[k], out.begin() + off[k] + len[k])});
        }
        done += n;
        frame += (uint64_t)nf;
    }
    close(fd);
    codec_private->resize(42);
    b200_flac_codec_private(E, total, codec_private->data(), 42);
    b200_flac_close(E);
    return 0;
}
