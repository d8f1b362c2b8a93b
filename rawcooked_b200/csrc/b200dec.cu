// C ABI of the B200 FFV1 decoder / checker (include/b200dec.h): handle management, device memory plan, launches.
#include "../../include/b200dec.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ffv1_dec.h"
#include "ffv1_host.h"

void b200_set_error(const std::string& msg);

namespace {
int dfail(int code, const std::string& msg) { b200_set_error(msg); return code; }
int dfail_cuda(cudaError_t e, const char* what) {
    b200_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA;
}
#define DCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return dfail_cuda(e_, #x); } while (0)
}  // namespace

struct b200_ffv1_dec {
    b200_ffv1_dec_cfg cfg;
    b200::Ffv1DecStream st;
    b200::DecArgs args;
    int max_frames = 0;
    size_t frame_bytes = 0;
    std::vector<void*> owned;
    // staging of the host entry points (grown on demand)
    uint8_t* d_pk = nullptr; size_t pk_cap = 0;
    uint8_t* d_frames = nullptr;        // max_frames payloads: decoded output or sources to compare with
    uint64_t *d_off = nullptr, *d_len = nullptr;
    uint64_t* h_meta = nullptr;         // pinned: [off n][len n]
    unsigned long long* h_mismatch = nullptr;   // pinned mirrors
    uint32_t* h_status = nullptr;
    unsigned long long* h_counters = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start, index done, decode done, results on the host
    uint64_t stats[8] = {0};
    bool pending = false;
    template <class T> cudaError_t alloc(T** p, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, n ? n : 16);
        if (e == cudaSuccess) { owned.push_back(q); *p = static_cast<T*>(q); }
        return e;
    }
};

extern "C" {

int b200_ffv1_parse_config_record(const uint8_t* record, size_t record_len, b200_ffv1_params* out) {
    if (!out) return dfail(B200_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof *out);
    b200::Ffv1DecStream S;
    std::string err;
    int r = b200::parse_config_record(record, record_len, &S, &err);
    out->crc_ok = S.crc_ok;
    if (r) return dfail(r, err);
    out->version = S.version; out->micro_version = S.micro; out->coder_type = S.coder_type_sent; out->colorspace_type = S.colorspace;
    out->bits_per_raw_sample = S.bits; out->chroma_planes = S.chroma_planes; out->log2_h_chroma_subsample = S.log2_h;
    out->log2_v_chroma_subsample = S.log2_v; out->alpha_plane = S.alpha; out->num_h_slices = S.num_h; out->num_v_slices = S.num_v;
    out->quant_table_set_count = S.nsets; out->ec = S.ec; out->intra = S.intra;
    for (int i = 0; i < 8; i++) out->context_count[i] = S.nctx[i];
    return 0;
}

int b200_ffv1_dec_open(const b200_ffv1_dec_cfg* cfg, const uint8_t* record, size_t record_len, b200_ffv1_dec** out) {
    if (!cfg || !out) return dfail(B200_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->max_frames < 1) return dfail(B200_ERR_INVALID, "max_frames must be >= 1");
    b200_ffv1_dec* D = new (std::nothrow) b200_ffv1_dec;
    if (!D) return dfail(B200_ERR_INVALID, "out of host memory");
    D->cfg = *cfg;
    std::string err;
    int r = b200::parse_config_record(record, record_len, &D->st, &err);
    if (r) { delete D; return dfail(r, err); }
    const b200::Ffv1DecStream& S = D->st;
    const int lbits = b200::layout_bits(cfg->layout);
    const char* why = nullptr;
    if (!lbits) why = "unsupported layout";
    else if (S.colorspace != 1 || !S.chroma_planes || S.log2_h || S.log2_v) why = "only RGB streams (colorspace_type 1) are decoded on the B200";
    else if (S.alpha) why = "streams with an alpha plane are not decoded on the B200";
    else if (S.coder_type_sent == 0) why = "Golomb-Rice streams (-coder 0) are not decoded on the B200";
    else if (!S.intra) why = "only intra streams (-g 1) are decoded on the B200";
    else if (S.bits != lbits) why = "the layout's bit depth differs from the stream's bits_per_raw_sample";
    else if (!cfg->width || !cfg->height || cfg->width > 65535 || cfg->height > 65535) why = "bad dimensions";
    else if ((uint32_t)S.num_h >= cfg->width || (uint32_t)S.num_v >= cfg->height) why = "FFV1-HEADER-num_h_slices:1";
    if (why) { delete D; return dfail(B200_ERR_INVALID, why); }

    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0) { delete D; return dfail(B200_ERR_NO_DEVICE, "no CUDA device: the B200 decoder has no CPU fallback"); }
    if (cfg->device < 0 || cfg->device >= ndev) { delete D; return dfail(B200_ERR_INVALID, "bad device ordinal"); }
    { cudaError_t e_ = cudaSetDevice(cfg->device); if (e_ != cudaSuccess) { delete D; return dfail_cuda(e_, "cudaSetDevice"); } }

    const int B = cfg->max_frames;
    D->max_frames = B;
    b200::DecArgs& A = D->args;
    std::memset(&A, 0, sizeof A);
    A.W = (int)cfg->width; A.H = (int)cfg->height; A.layout = cfg->layout; A.bits = S.bits; A.bits_max = S.bits + 1;
    A.swap_bg = S.bits > 8 && S.bits < 16;
    A.num_h = S.num_h; A.num_v = S.num_v; A.nslices = S.num_h * S.num_v; A.ec = S.ec; A.tail = S.ec ? 8 : 3; A.nsets = S.nsets;
    int spw = cfg->slices_per_warp;
    if (!spw) { const char* e = std::getenv("B200_DEC_SPW"); spw = e ? std::atoi(e) : 0; }
    if (spw < 0 || spw > 32) { delete D; return dfail(B200_ERR_INVALID, "slices_per_warp must be 1..32"); }
    A.spw = spw;     // 0: chosen per call from the number of slices in flight
    A.maxctx = 0;
    for (int i = 0; i < S.nsets; i++) { A.nctx[i] = S.nctx[i]; if (S.nctx[i] > A.maxctx) A.maxctx = S.nctx[i]; }
    A.row_bytes = (uint32_t)b200::layout_row_bytes(cfg->width, cfg->layout);
    A.frame_bytes = (size_t)A.row_bytes * cfg->height;
    D->frame_bytes = A.frame_bytes;
    // widest slice of the grid (boundaries as FFV1_Slice.cpp:152-155)
    int wmax = 0;
    for (int sx = 0; sx < S.num_h; sx++) {
        const int x0 = (int)((uint64_t)sx * cfg->width / S.num_h), x1 = (int)((uint64_t)(sx + 1) * cfg->width / S.num_h);
        if (x1 - x0 > wmax) wmax = x1 - x0;
    }
    // a slice header may declare a slice several grid cells wide: such a slice is refused (BAD_HEADER) when it exceeds wpad
    A.wpad = (wmax + 3) & ~3;

    const size_t nsl = (size_t)B * A.nslices;
    int16_t* d_q = nullptr; uint8_t* d_t = nullptr; uint32_t* d_c = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(D->alloc(&d_q, S.qtab.size() * 2));
    ok(D->alloc(&d_t, 512));
    ok(D->alloc(&d_c, 1024));
    ok(D->alloc(&A.sl_off, nsl * 8));
    ok(D->alloc(&A.sl_size, nsl * 4));
    ok(D->alloc(&A.states, nsl * 2 * (size_t)A.maxctx * 32));
    ok(D->alloc(&A.lines, nsl * 9 * (size_t)A.wpad * 4));
    ok(D->alloc(&A.mismatch, (size_t)B * 8));
    ok(D->alloc(&A.status, (size_t)B * 4));
    ok(D->alloc(&A.counters, 16));
    ok(D->alloc(&D->d_off, (size_t)B * 8));
    ok(D->alloc(&D->d_len, (size_t)B * 8));
    if (e == cudaSuccess) {
        uint8_t trans[512];
        std::memcpy(trans, S.zero_state, 256);
        std::memcpy(trans + 256, S.one_state, 256);
        ok(cudaMemcpy(d_q, S.qtab.data(), S.qtab.size() * 2, cudaMemcpyHostToDevice));
        ok(cudaMemcpy(d_t, trans, 512, cudaMemcpyHostToDevice));
        ok(cudaMemcpy(d_c, b200::crc32_mpeg_table(), 1024, cudaMemcpyHostToDevice));
        ok(cudaMallocHost(reinterpret_cast<void**>(&D->h_meta), (size_t)B * 16));
        ok(cudaMallocHost(reinterpret_cast<void**>(&D->h_mismatch), (size_t)B * 8));
        ok(cudaMallocHost(reinterpret_cast<void**>(&D->h_status), (size_t)B * 4));
        ok(cudaMallocHost(reinterpret_cast<void**>(&D->h_counters), 16));
        ok(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
        for (auto& v : D->ev) ok(cudaEventCreate(&v));
    }
    if (e != cudaSuccess) { b200_ffv1_dec_close(D); return dfail_cuda(e, "decoder allocation"); }
    A.qtab = d_q; A.trans = d_t; A.crc_table = d_c;
    A.pkt_off = D->d_off; A.pkt_len = D->d_len;
    *out = D;
    return 0;
}

void b200_ffv1_dec_close(b200_ffv1_dec* D) {
    if (!D) return;
    cudaSetDevice(D->cfg.device);
    if (D->stream) { cudaStreamSynchronize(D->stream); cudaStreamDestroy(D->stream); }
    for (auto v : D->ev) if (v) cudaEventDestroy(v);
    for (void* p : D->owned) cudaFree(p);
    if (D->d_pk) cudaFree(D->d_pk);
    if (D->d_frames) cudaFree(D->d_frames);
    if (D->h_meta) cudaFreeHost(D->h_meta);
    if (D->h_mismatch) cudaFreeHost(D->h_mismatch);
    if (D->h_status) cudaFreeHost(D->h_status);
    if (D->h_counters) cudaFreeHost(D->h_counters);
    delete D;
}

int b200_ffv1_decode_device(b200_ffv1_dec* D, const void* d_packets, const size_t* pkt_off, const size_t* pkt_len, int32_t n,
                            void* d_out, const void* d_sources, void* stream) {
    if (!D || !d_packets || !pkt_off || !pkt_len) return dfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > D->max_frames) return dfail(B200_ERR_INVALID, "n_frames out of range");
    DCU(cudaSetDevice(D->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    b200::DecArgs A = D->args;
    for (int i = 0; i < n; i++) { D->h_meta[i] = pkt_off[i]; D->h_meta[n + i] = pkt_len[i]; }
    DCU(cudaMemcpyAsync(D->d_off, D->h_meta, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    DCU(cudaMemcpyAsync(D->d_len, D->h_meta + n, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    if (!A.spw) {
        // a lane is latency-bound (one dependent instruction every ~4.6 cycles), so the more warps in flight the better, up to
        // one wave: 148 SMs x 16 warps at 128 registers per thread. Beyond that, slices share warps.
        const int total = n * A.nslices, wave = 148 * 16;
        A.spw = (total + wave - 1) / wave;
        if (A.spw > 8) A.spw = 8;
    }
    A.packets = static_cast<const uint8_t*>(d_packets);
    A.out = static_cast<uint8_t*>(d_out);
    A.cmp = static_cast<const uint8_t*>(d_sources);
    DCU(cudaEventRecord(D->ev[0], s));
    // every frame is a keyframe: all context states start at 128 (FFV1_Coder_RangeCoder.cpp:25-48; states are not coded)
    DCU(cudaMemsetAsync(A.states, 0x80, (size_t)n * A.nslices * 2 * (size_t)A.maxctx * 32, s));
    DCU(cudaMemsetAsync(A.counters, 0, 16, s));
    // row padding is written as zeros; 12-bit packed words shared by two slices are OR-ed together
    if (A.out && (A.layout == B200_DPX_RGB_12_PACKED_BE || A.layout == B200_DPX_RGB_8 || A.layout == B200_DPX_RGB_16_LE || A.layout == B200_DPX_RGB_16_BE))
        DCU(cudaMemsetAsync(A.out, 0, (size_t)n * A.frame_bytes, s));
    DCU(b200::launch_dec_index(A, n, s));
    DCU(cudaEventRecord(D->ev[1], s));
    DCU(b200::launch_decode(A, n, s));
    DCU(cudaEventRecord(D->ev[2], s));
    DCU(cudaMemcpyAsync(D->h_mismatch, A.mismatch, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    DCU(cudaMemcpyAsync(D->h_status, A.status, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    DCU(cudaMemcpyAsync(D->h_counters, A.counters, 16, cudaMemcpyDeviceToHost, s));
    DCU(cudaEventRecord(D->ev[3], s));
    D->pending = true;
    return 0;
}

int b200_ffv1_dec_result(b200_ffv1_dec* D, uint64_t* mismatch, uint32_t* status, int32_t n) {
    if (!D) return dfail(B200_ERR_INVALID, "null argument");
    if (!D->pending) return dfail(B200_ERR_INVALID, "no decode call to collect");
    if (n < 1 || n > D->max_frames) return dfail(B200_ERR_INVALID, "n_frames out of range");
    DCU(cudaSetDevice(D->cfg.device));
    DCU(cudaEventSynchronize(D->ev[3]));
    for (int i = 0; i < n; i++) {
        if (mismatch) mismatch[i] = D->h_mismatch[i];
        if (status) status[i] = D->h_status[i];
    }
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, D->ev[0], D->ev[1]);
    cudaEventElapsedTime(&b, D->ev[1], D->ev[2]);
    D->stats[0] = (uint64_t)(a * 1000.0f); D->stats[1] = (uint64_t)(b * 1000.0f); D->stats[2] = D->stats[0] + D->stats[1];
    D->stats[3] = D->h_counters[0]; D->stats[4] = D->h_counters[1];
    D->pending = false;
    return 0;
}

static int run_host(b200_ffv1_dec* D, const uint8_t* const* packets, const size_t* packet_len, int32_t n,
                    uint8_t* const* frames, const uint8_t* const* sources, uint64_t* mismatch, uint32_t* status) {
    if (!D || !packets || !packet_len) return dfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > D->max_frames) return dfail(B200_ERR_INVALID, "n_frames out of range");
    DCU(cudaSetDevice(D->cfg.device));
    std::vector<size_t> off(n), len(n);
    size_t total = 0;
    for (int i = 0; i < n; i++) { off[i] = total; len[i] = packet_len[i]; total += (packet_len[i] + 15) & ~(size_t)15; }
    if (total + 16 > D->pk_cap) {
        if (D->d_pk) cudaFree(D->d_pk);
        D->d_pk = nullptr; D->pk_cap = 0;
        DCU(cudaMalloc(reinterpret_cast<void**>(&D->d_pk), total + 16));
        D->pk_cap = total + 16;
    }
    if (!D->d_frames) DCU(cudaMalloc(reinterpret_cast<void**>(&D->d_frames), (size_t)D->max_frames * D->frame_bytes + 16));
    cudaStream_t s = D->stream;
    for (int i = 0; i < n; i++) DCU(cudaMemcpyAsync(D->d_pk + off[i], packets[i], len[i], cudaMemcpyHostToDevice, s));
    if (sources)
        for (int i = 0; i < n; i++) DCU(cudaMemcpyAsync(D->d_frames + (size_t)i * D->frame_bytes, sources[i], D->frame_bytes, cudaMemcpyHostToDevice, s));
    int r = b200_ffv1_decode_device(D, D->d_pk, off.data(), len.data(), n, sources ? nullptr : D->d_frames, sources ? D->d_frames : nullptr, s);
    if (r) return r;
    if (frames)
        for (int i = 0; i < n; i++) DCU(cudaMemcpyAsync(frames[i], D->d_frames + (size_t)i * D->frame_bytes, D->frame_bytes, cudaMemcpyDeviceToHost, s));
    DCU(cudaStreamSynchronize(s));
    return b200_ffv1_dec_result(D, mismatch, status, n);
}

int b200_ffv1_decode_host(b200_ffv1_dec* D, const uint8_t* const* packets, const size_t* packet_len, int32_t n,
                          uint8_t* const* frames, uint32_t* status) {
    if (!frames) return dfail(B200_ERR_INVALID, "null argument");
    return run_host(D, packets, packet_len, n, frames, nullptr, nullptr, status);
}

int b200_ffv1_check_host(b200_ffv1_dec* D, const uint8_t* const* packets, const size_t* packet_len, int32_t n,
                         const uint8_t* const* sources, uint64_t* mismatch, uint32_t* status) {
    if (!sources) return dfail(B200_ERR_INVALID, "null argument");
    return run_host(D, packets, packet_len, n, nullptr, sources, mismatch, status);
}

int b200_ffv1_dec_stats(const b200_ffv1_dec* D, uint64_t stats[8]) {
    if (!D || !stats) return dfail(B200_ERR_INVALID, "null argument");
    std::memcpy(stats, D->stats, sizeof D->stats);
    return 0;
}

}  // extern "C"
