// C ABI of the B200 FFV1 encoder (include/b200enc.h): handle management, device memory plan, band loop.
#include "../../include/b200enc.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ffv1_host.h"
#include "ffv1_kernels.cuh"
#include "sm_partition.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, #x); } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t n, std::vector<void*>& owned) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n ? n : 16);
    if (e == cudaSuccess) { owned.push_back(q); *p = static_cast<T*>(q); }
    return e;
}

}  // namespace

void b200_set_error(const std::string& msg) { g_err = msg; }

struct b200_ffv1_enc {
    b200_ffv1_cfg cfg;
    b200::Ffv1Stream st;
    b200::EncArgs args;             // band buffers of parity 0
    static constexpr int kPar = 3;  // most band-buffer sets
    int kpar = 2;                   // band-buffer sets in rotation (B200_KPAR=3 with B200_MODEL_RESERVE=36 measured +15 % at
                                    // B = 128, but the optimum moves with B and with where the CTAs land: DESIGN.md §4)
    b200::EncArgs argsN[kPar];      // [0] unused (= args); [p] same as args with the band buffers of set p
    cudaStream_t sm = nullptr, sr = nullptr, se = nullptr;   // model / range / emit streams
    // SM partition (green contexts): k_range on SMs of its own, k_model on the rest; k_emit either on a third part
    // (emit_mode 0), behind k_model on the model stream (1, default) or behind k_range on the range stream (2)
    b200::SmPartition part;
    int emit_mode = 0;
    bool scratch_dirty = true;                  // the slice scratch needs a full memset before the next encode
    bool own_streams = true, own_se = false;    // sm/sr (and se) were made by cudaStreamCreate, not by the partition
    cudaStream_t sc = nullptr;                               // host-to-device copies of the host entry point
    std::vector<cudaEvent_t> ev_h2d;                         // [band] rows of the band have arrived
    cudaEvent_t ev_start = nullptr, ev_model[kPar] = {}, ev_range[kPar] = {}, ev_emit[kPar] = {};
    cudaEvent_t ev_done_m = nullptr, ev_done_e = nullptr;
    int max_frames = 0;
    std::vector<void*> owned;       // device allocations
    uint8_t* d_in = nullptr;        // staging for the host entry point (allocated on first use)
    uint8_t* d_in2 = nullptr;       // second staging buffer: the frames of the next batch cross PCIe while this one is coded
    cudaEvent_t ev_prefetch = nullptr;
    int prefetched_set = -1;        // b200_ffv1_prefetch_host has put a batch into this staging buffer (-1: none)
    const uint8_t* prefetched_first = nullptr;
    int prefetched_n = 0;
    size_t max_packet = 0;
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> tev;   // timing mode: 4 events per band + 2 around the pack kernels
    std::vector<cudaEvent_t> trace; // B200_TRACE: [band][model start, model end, range start, range end, emit start, emit end] on the kernels' own streams
    bool trace_pending = false;
    bool timed_pending = false;
    uint64_t stats[8] = {0};
    // Results of an encode call (packet arena, slice / frame tables, flags). Two sets: while the packets of one batch
    // cross PCIe, the next batch submitted with b200_ffv1_submit_host is already coding into the other set. Set 1 is
    // allocated the first time two host batches are in flight; the device entry point always uses set 0.
    struct ResultSet {
        uint8_t* arena = nullptr;
        uint32_t* slice_size = nullptr;
        uint64_t *slice_off = nullptr, *frame_off = nullptr, *frame_len = nullptr;
        uint32_t* flags = nullptr;
        uint32_t* h_flags = nullptr;    // pinned mirror of flags
        std::vector<uint64_t> h_off, h_len;
        cudaEvent_t done = nullptr;
        int n_frames = 0;
    } rs[2];
    int fifo[2] = {0, 0};           // result sets submitted and not yet collected, oldest first
    int fifo_n = 0;
    bool host_mode = false;         // batches come from b200_ffv1_submit_host (queue semantics) / from b200_ffv1_encode_device (the last call stays collectable)
    cudaStream_t sh = nullptr, sf = nullptr;   // stream of the host entry points / of the device-to-host fetches
};

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
uint32_t b200_version(void) { return (0u << 16) | (1u << 8) | 0u; }

int b200_ffv1_slice_grid(uint32_t width, uint32_t height, int32_t slices, int32_t* num_h, int32_t* num_v) {
    int h = 0, v = 0;
    // the grid search does not depend on the bit depth except through a size cap that only matters above 8K
    int r = b200::slice_grid(width, height, slices, 16, &h, &v);
    if (r) return fail(B200_ERR_INVALID, "no slice grid for this -slices value");
    if (num_h) *num_h = h;
    if (num_v) *num_v = v;
    return 0;
}

size_t b200_ffv1_frame_bytes(uint32_t width, uint32_t height, int32_t layout) {
    return b200::layout_row_bytes(width, layout) * height;
}

int b200_ffv1_open(const b200_ffv1_cfg* cfg, b200_ffv1_enc** out) {
    if (!cfg || !out) return fail(B200_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->coder != 1) return fail(B200_ERR_INVALID, "only -coder 1 (range coder) is implemented");
    if (cfg->max_frames < 1) return fail(B200_ERR_INVALID, "max_frames must be >= 1");
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0) return fail(B200_ERR_NO_DEVICE, "no CUDA device: the B200 encoder has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(B200_ERR_INVALID, "bad device ordinal");
    CU(cudaSetDevice(cfg->device));

    b200_ffv1_enc* E = new (std::nothrow) b200_ffv1_enc;
    if (!E) return fail(B200_ERR_INVALID, "out of host memory");
    E->cfg = *cfg;
    const char* err = "";
    int r = b200::build_stream(cfg->width, cfg->height, cfg->layout, cfg->slices, cfg->context, cfg->slicecrc, &E->st, &err);
    if (r) { delete E; return fail(r, err); }
    const b200::Ffv1Stream& S = E->st;
    const int ns = (int)S.slices.size();
    const int B = cfg->max_frames;
    E->max_frames = B;

    int wmax = 0, hmax = 0;
    for (const auto& g : S.slices) { wmax = g.w > wmax ? g.w : wmax; hmax = g.h > hmax ? g.h : hmax; }
    wmax = (wmax + 3) & ~3;

    b200::EncArgs& A = E->args;
    std::memset(&A, 0, sizeof A);
    A.W = (int)S.width; A.H = (int)S.height; A.layout = S.layout; A.bits = S.bits; A.sbits = S.sbits; A.swap_bg = S.swap_bg;
    A.nslices = ns; A.nctx = S.nctx; A.is5 = S.is5; A.ec = S.ec;
    A.row_bytes = (uint32_t)S.row_bytes; A.frame_bytes = S.frame_bytes;
    A.band_rows = 16;
    if (const char* e = getenv("B200_BAND_ROWS")) { int v = atoi(e); if (v >= 1 && v <= 4096) A.band_rows = v; }
    if (A.band_rows > 64) A.band_rows = 64;
    if (A.band_rows > hmax) A.band_rows = hmax;
    A.nbands = (hmax + A.band_rows - 1) / A.band_rows;
    A.wmax = wmax; A.hmax = hmax;
    A.sstride = S.sbits <= 9 ? 28 : 32;

    if (wmax > 2048) { delete E; return fail(B200_ERR_INVALID, "slice wider than 2048 pixels: use more slices"); }
    // The model kernel keeps the whole context-state table of one plane-set in shared memory; what is left (minus a
    // reserve that lets k_range CTAs run on the same SM) stages the records of one plane-row. Rows whose records do not fit are
    // coded in up to kMaxSeg column segments.
    int dev_smem = 0;
    cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const int maxbins = 2 * S.sbits + 1;                        // bins of the largest symbol: 2e+3 with e = sbits-1
    const size_t row_need = (((size_t)wmax * maxbins + b200::kMaxHeaderBins) + 127) & ~(size_t)127;
    const size_t chunk_need = 32 * (size_t)maxbins + b200::kMaxHeaderBins;
    size_t reserve = 0;      // k_range has SMs of its own (SM partition); nothing needs to co-reside with k_model
    if (const char* e = getenv("B200_SMEM_RESERVE")) reserve = (size_t)atoi(e);
    int best = -1, best_nseg = 0;
    size_t best_cap = 0;
    {
        const size_t fixed = b200::model_smem_fixed(S.nctx, A.sstride, wmax, 2);
        if (fixed + reserve + 2 * (chunk_need + 256) <= (size_t)dev_smem) {
            size_t capk = (((size_t)dev_smem - reserve - fixed) / 2) & ~(size_t)127;
            if (capk > row_need) capk = row_need;
            if (capk > 32640) capk = 32640;             // record offsets inside the stage travel as 16-bit byte offsets
            best = 0;
            best_cap = capk;
            best_nseg = capk >= row_need ? 1 : (int)(((size_t)wmax * maxbins + (capk - chunk_need) - 1) / (capk - chunk_need));
        }
    }
    if (best < 0 || best_nseg > b200::kMaxSeg) {
        delete E;
        return fail(B200_ERR_INVALID, "slice too wide for the shared-memory context model: use more slices");
    }
    A.stage_cap = (int32_t)best_cap;
    A.nseg = best_nseg;
    if (A.band_rows * 3 * A.nseg > 4096) { delete E; return fail(B200_ERR_INVALID, "band too large"); }
    cudaError_t ce = b200::configure_kernels(A);
    if (ce != cudaSuccess) { delete E; return fail_cuda(ce, "configure_kernels"); }

    // records of one band of one slice: every plane-row segment is padded to whole 128-record blocks
    A.capY = (size_t)A.band_rows * (row_need + 128 * (size_t)A.nseg) + 128;
    A.capC = A.capY * 2;
    size_t samples = (size_t)wmax * hmax * 3;
    // slice byte streams: FFmpeg's own worst-case bound per sample (white noise on small 16-bit slices really reaches 1.7 x raw)
    A.slice_cap = ((samples * (2 * S.bits + 5) / 8 + 1024 + 15) & ~(size_t)15);
    E->max_packet = A.slice_cap * ns;
    A.arena_cap = E->max_packet * B;

    auto& own = E->owned;
    std::vector<uint16_t> hb((size_t)ns * b200::kMaxHeaderBins, 0);
    std::vector<int32_t> hc(ns, 0);
    for (int i = 0; i < ns; i++) {
        if (S.header_bins[i].size() > (size_t)b200::kMaxHeaderBins) { delete E; return fail(B200_ERR_INVALID, "slice header too long"); }
        hc[i] = (int32_t)S.header_bins[i].size();
        for (size_t k = 0; k < S.header_bins[i].size(); k++) {      // (state | bit << 8) -> coder record (sp - 1) | bit << 8
            const uint16_t r = S.header_bins[i][k];
            const uint32_t st = r & 255u, bit = r >> 8;
            hb[(size_t)i * b200::kMaxHeaderBins + k] = (uint16_t)(((bit ? st : 256u - st) - 1u) | (bit << 8));
        }
    }
    uint8_t t1q[256];                                               // one_state indexed by q = sp - 1
    for (int i = 0; i < 255; i++) t1q[i] = S.one_state[i + 1];
    t1q[255] = 0;
    uint8_t tpow[5][256];                                           // one_state iterated 2^i times (runs of zero residuals)
    for (int i = 0; i < 256; i++) tpow[0][i] = i ? S.one_state[i] : 0;
    for (int k = 1; k < 5; k++)
        for (int i = 0; i < 256; i++) tpow[k][i] = tpow[k - 1][tpow[k - 1][i]];

    b200::SliceGeom* d_geom; int16_t* d_qtab; uint8_t* d_trans; uint16_t* d_hb; int32_t* d_hc; uint32_t* d_crc;
#define ALLOC(p, n) do { cudaError_t e_ = dalloc(&(p), (n), own); if (e_ != cudaSuccess) { int rc_ = fail_cuda(e_, "cudaMalloc " #p); b200_ffv1_close(E); return rc_; } } while (0)
    ALLOC(d_geom, sizeof(b200::SliceGeom) * ns);
    ALLOC(d_qtab, sizeof S.qtab);
    ALLOC(d_trans, 256);
    uint8_t* d_tpow;
    ALLOC(d_tpow, sizeof tpow);
    ALLOC(d_hb, hb.size() * 2);
    ALLOC(d_hc, hc.size() * 4);
    ALLOC(d_crc, 1024);
    ALLOC(A.state_save, (size_t)B * ns * 2 * (((size_t)S.nctx * A.sstride + 15) & ~(size_t)15));
    ALLOC(A.qY, (size_t)B * ns * A.capY);
    ALLOC(A.qC, (size_t)B * ns * A.capC);
    ALLOC(A.bY, (size_t)B * ns * (A.capY >> 3));
    ALLOC(A.bC, (size_t)B * ns * (A.capC >> 3));
    ALLOC(A.rowcnt, (size_t)B * ns * A.band_rows * 3 * A.nseg * 4);
    ALLOC(A.ckptY, (size_t)B * ns * (A.capY >> 6) * 8);
    ALLOC(A.ckptC, (size_t)B * ns * (A.capC >> 6) * 8);
    ALLOC(A.used, (size_t)B * ns * 2 * 4);
    if (const char* e = getenv("B200_KPAR")) E->kpar = atoi(e) == 3 ? 3 : 2;
    for (int pz = 1; pz < E->kpar; pz++) {
        b200::EncArgs& P = E->argsN[pz];
        ALLOC(P.qY, (size_t)B * ns * A.capY);
        ALLOC(P.qC, (size_t)B * ns * A.capC);
        ALLOC(P.bY, (size_t)B * ns * (A.capY >> 3));
        ALLOC(P.bC, (size_t)B * ns * (A.capC >> 3));
        ALLOC(P.rowcnt, (size_t)B * ns * A.band_rows * 3 * A.nseg * 4);
        ALLOC(P.ckptY, (size_t)B * ns * (A.capY >> 6) * 8);
        ALLOC(P.ckptC, (size_t)B * ns * (A.capC >> 6) * 8);
        ALLOC(P.used, (size_t)B * ns * 2 * 4);
    }
    ALLOC(A.cstate, (size_t)B * ns * sizeof(b200::CoderState));
    ALLOC(A.scratch, (size_t)B * ns * A.slice_cap);
    ALLOC(A.slice_size, (size_t)B * ns * 4);
    ALLOC(A.slice_off, (size_t)B * ns * 8);
    ALLOC(A.frame_off, (size_t)B * 8);
    ALLOC(A.frame_len, (size_t)B * 8);
    ALLOC(A.arena, A.arena_cap);
    ALLOC(A.flags, 256);
    ALLOC(A.work_ctr, (size_t)A.nbands * 4);
    {
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device);
        int reserve = 0;
        if (const char* e = getenv("B200_MODEL_RESERVE")) reserve = atoi(e);
        A.model_ctas = nsm - reserve > 1 ? nsm - reserve : 1;
        A.range_smem = 28 * 1024;
        if (const char* e = getenv("B200_RANGE_SMEM_KB")) A.range_smem = atoi(e) * 1024;
        if (A.range_smem > 48 * 1024) A.range_smem = 48 * 1024;
    }
#undef ALLOC
    cudaMemcpy(d_geom, S.slices.data(), sizeof(b200::SliceGeom) * ns, cudaMemcpyHostToDevice);
    cudaMemcpy(d_qtab, S.qtab, sizeof S.qtab, cudaMemcpyHostToDevice);
    cudaMemcpy(d_trans, t1q, 256, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tpow, tpow, sizeof tpow, cudaMemcpyHostToDevice);
    cudaMemcpy(d_hb, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_hc, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_crc, b200::crc32_mpeg_table(), 1024, cudaMemcpyHostToDevice);
    A.geom = d_geom; A.qtab = d_qtab; A.t1q = d_trans; A.tpow = d_tpow; A.hdr_bins = d_hb; A.hdr_cnt = d_hc; A.crc_table = d_crc;
    for (int pz = 1; pz < E->kpar; pz++) {
        const b200::EncArgs P = E->argsN[pz];
        E->argsN[pz] = A;
        b200::EncArgs& Q = E->argsN[pz];
        Q.qY = P.qY; Q.qC = P.qC; Q.bY = P.bY; Q.bC = P.bC; Q.rowcnt = P.rowcnt; Q.ckptY = P.ckptY; Q.ckptC = P.ckptC; Q.used = P.used;
    }
    // equal priorities: with prioritised streams the device preempts k_model's CTAs (227 KB of state each) whenever
    // k_range / k_emit become runnable, which costs more than it gains (measured: 28.5 ms per band against 21 ms)
    {
        // Partition: B200_RANGE_SMS SMs (default 24) to k_range, optionally B200_EMIT_SMS to k_emit, the rest to k_model.
        // B200_NO_PARTITION=1 (or a driver without green contexts) falls back to three plain streams on the whole device.
        // k_range: one lane per (frame, slice), CTAs of 128 lanes, two CTAs per SM (two warps per scheduler cost a fifth of
        // their speed and halve the SMs taken from k_model); the driver hands out SMs in groups of 8
        const int range_ctas = (B * ns + 127) / 128;
        int range_sms = (((range_ctas + 1) / 2) + 7) / 8 * 8, emit_sms = 0;
        if (range_sms > 32) range_sms = 32;
        if (const char* e = getenv("B200_RANGE_SMS")) range_sms = atoi(e);
        if (const char* e = getenv("B200_EMIT_SMS")) emit_sms = atoi(e);
        E->emit_mode = emit_sms > 0 ? 0 : 1;
        if (const char* e = getenv("B200_EMIT_MODE")) E->emit_mode = atoi(e);
        if (emit_sms > 0 && E->emit_mode != 0) emit_sms = 0;
        bool parted = false;
        if (!getenv("B200_NO_PARTITION") && range_sms > 0) {
            std::vector<int> want;
            want.push_back(range_sms);
            if (emit_sms > 0) want.push_back(emit_sms);
            want.push_back(0);
            if (E->part.create(cfg->device, want)) {
                const int last = E->part.parts() - 1;
                E->sr = E->part.stream(0);
                E->se = emit_sms > 0 ? E->part.stream(1) : nullptr;
                E->sm = E->part.stream(last);
                parted = E->sr && E->sm && (emit_sms == 0 || E->se);
                if (parted) {
                    A.model_ctas = E->part.sm_count(last);
                    A.range_smem = 0;                    // the record rings alone (80 KB): two k_range CTAs per SM
                    if (const char* e = getenv("B200_RANGE_CTAS_PER_SM")) { if (atoi(e) == 1) A.range_smem = 120 * 1024; }
                    A.range_sms = E->part.sm_count(0);
                    for (int pz = 1; pz < E->kpar; pz++) {
                        E->argsN[pz].model_ctas = A.model_ctas; E->argsN[pz].range_smem = A.range_smem; E->argsN[pz].range_sms = A.range_sms;
                    }
                }
            } else if (getenv("B200_VERBOSE")) {
                fprintf(stderr, "b200enc: no SM partition (%s)\n", E->part.why().c_str());
            }
        }
        if (!parted) {
            E->sm = E->sr = E->se = nullptr;
            cudaStreamCreateWithFlags(&E->sm, cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&E->sr, cudaStreamNonBlocking);
            if (E->emit_mode != 0 && !getenv("B200_EMIT_MODE")) E->emit_mode = 0;
        }
        E->own_streams = !parted;
        if (E->emit_mode == 0 && !E->se) { cudaStreamCreateWithFlags(&E->se, cudaStreamNonBlocking); E->own_se = true; }
        if (getenv("B200_VERBOSE"))
            fprintf(stderr, "b200enc: %s, k_model grid %d, emit mode %d\n", parted ? "SM partition on" : "whole device", A.model_ctas, E->emit_mode);
    }
    cudaStreamCreateWithFlags(&E->sc, cudaStreamNonBlocking);
    E->ev_h2d.resize(A.nbands);
    for (auto& ev : E->ev_h2d) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (cudaEvent_t* ev : {&E->ev_start, &E->ev_done_m, &E->ev_done_e, &E->ev_prefetch}) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    for (int pz = 0; pz < b200_ffv1_enc::kPar; pz++)
        for (cudaEvent_t* ev : {&E->ev_model[pz], &E->ev_range[pz], &E->ev_emit[pz]}) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    {
        b200_ffv1_enc::ResultSet& R = E->rs[0];
        R.arena = A.arena; R.slice_size = A.slice_size; R.slice_off = A.slice_off; R.frame_off = A.frame_off; R.frame_len = A.frame_len;
        R.flags = A.flags;
    }
    cudaError_t e2 = cudaSuccess;
    for (auto& R : E->rs) {
        if (e2 == cudaSuccess) e2 = cudaHostAlloc((void**)&R.h_flags, 256, cudaHostAllocDefault);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&R.done, cudaEventDisableTiming);
        R.h_off.resize(B); R.h_len.resize(B);
    }
    if (e2 != cudaSuccess) { int rc = fail_cuda(e2, "cudaHostAlloc"); b200_ffv1_close(E); return rc; }
    cudaStreamCreateWithFlags(&E->sh, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&E->sf, cudaStreamNonBlocking);
    for (auto& ev : E->ev) cudaEventCreate(&ev);
    e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { int rc = fail_cuda(e2, "init"); b200_ffv1_close(E); return rc; }
    *out = E;
    return 0;
}

void b200_ffv1_close(b200_ffv1_enc* E) {
    if (!E) return;
    cudaSetDevice(E->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : E->owned) cudaFree(p);
    if (E->d_in) cudaFree(E->d_in);
    if (E->d_in2) cudaFree(E->d_in2);
    if (E->ev_prefetch) cudaEventDestroy(E->ev_prefetch);
    for (auto& R : E->rs) {
        if (R.h_flags) cudaFreeHost(R.h_flags);
        if (R.done) cudaEventDestroy(R.done);
    }
    {   // set 1 is allocated outside `owned`
        b200_ffv1_enc::ResultSet& R = E->rs[1];
        for (void* q : {(void*)R.arena, (void*)R.slice_size, (void*)R.slice_off, (void*)R.frame_off, (void*)R.frame_len, (void*)R.flags})
            if (q) cudaFree(q);
    }
    for (cudaStream_t st : {E->sh, E->sf}) if (st) cudaStreamDestroy(st);
    for (auto& ev : E->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : E->tev) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {E->ev_start, E->ev_done_m, E->ev_done_e}) if (ev) cudaEventDestroy(ev);
    for (int pz = 0; pz < b200_ffv1_enc::kPar; pz++)
        for (cudaEvent_t ev : {E->ev_model[pz], E->ev_range[pz], E->ev_emit[pz]}) if (ev) cudaEventDestroy(ev);
    for (auto& ev : E->ev_h2d) if (ev) cudaEventDestroy(ev);
    if (E->own_streams) for (cudaStream_t st : {E->sm, E->sr}) if (st) cudaStreamDestroy(st);
    if (E->own_se && E->se) cudaStreamDestroy(E->se);
    if (E->sc) cudaStreamDestroy(E->sc);
    delete E;
}

size_t b200_ffv1_config_record(const b200_ffv1_enc* E, uint8_t* out, size_t cap) {
    if (!E) return 0;
    size_t n = E->st.config_record.size();
    if (out && cap) std::memcpy(out, E->st.config_record.data(), n < cap ? n : cap);
    return n;
}

size_t b200_ffv1_config_record_for(const b200_ffv1_cfg* cfg, uint8_t* out, size_t cap) {
    if (!cfg) { fail(B200_ERR_INVALID, "null argument"); return 0; }
    if (cfg->coder != 1) { fail(B200_ERR_INVALID, "only -coder 1 (range coder) is implemented"); return 0; }
    b200::Ffv1Stream st;
    const char* err = "";
    const int r = b200::build_stream(cfg->width, cfg->height, cfg->layout, cfg->slices, cfg->context, cfg->slicecrc, &st, &err);
    if (r) { fail(r, err); return 0; }
    const size_t n = st.config_record.size();
    if (out && cap) std::memcpy(out, st.config_record.data(), n < cap ? n : cap);
    return n;
}

size_t b200_ffv1_max_packet_bytes(const b200_ffv1_enc* E) { return E ? E->max_packet : 0; }

int b200_ffv1_set_timing(b200_ffv1_enc* E, int32_t enabled) {
    if (!E) return fail(B200_ERR_INVALID, "null encoder");
    E->timing = enabled != 0;
    return 0;
}

// host_frames != nullptr: the payloads are still in host memory; they are copied to d_frames band by band on the copy stream,
// each band's rows just ahead of the k_model launch that needs them, so that the transfer hides behind the kernels
// whole frames of a batch into a staging buffer, on the copy stream, behind the last kernels that read that buffer
static int prefetch_copies(b200_ffv1_enc* E, uint8_t* d_buf, const uint8_t* const* host_frames, int32_t n_frames, int set) {
    // the buffer was last read by the batch that used this result set: its `done` event covers every kernel of it
    CU(cudaStreamWaitEvent(E->sc, E->rs[set].done, 0));
    for (int i = 0; i < n_frames; i++)
        CU(cudaMemcpyAsync(d_buf + (size_t)i * E->st.frame_bytes, host_frames[i], E->st.frame_bytes, cudaMemcpyHostToDevice, E->sc));
    CU(cudaEventRecord(E->ev_prefetch, E->sc));
    return 0;
}

// prefetch (a batch is already in flight): the frames go to the staging buffer whole and at once, on the copy stream, while
// the batch ahead is still being coded from the other staging buffer; the first k_model launch waits for all of them. Large
// copies run at the full PCIe rate (the band-sized pieces of 368 KB reach ~16 GB/s next to the packet download, which made
// the end-to-end path transfer-bound on 4K 16-bit), and nothing of the transfer is left between two batches.
static int encode_impl(b200_ffv1_enc* E, const void* d_frames, int32_t n_frames, cudaStream_t s, const uint8_t* const* host_frames, int set,
                       bool prefetch = false) {
    CU(cudaSetDevice(E->cfg.device));
    b200_ffv1_enc::ResultSet& R = E->rs[set];
    if (!R.arena) {     // second result set, first use
        const b200::EncArgs& A0 = E->args;
        const size_t B = (size_t)E->max_frames, ns = (size_t)A0.nslices;
        CU(cudaMalloc((void**)&R.arena, A0.arena_cap));
        CU(cudaMalloc((void**)&R.slice_size, B * ns * 4));
        CU(cudaMalloc((void**)&R.slice_off, B * ns * 8));
        CU(cudaMalloc((void**)&R.frame_off, B * 8));
        CU(cudaMalloc((void**)&R.frame_len, B * 8));
        CU(cudaMalloc((void**)&R.flags, 256));
    }
    constexpr int kPar = b200_ffv1_enc::kPar;
    b200::EncArgs A[kPar];
    A[0] = E->args;
    for (int pz = 1; pz < E->kpar; pz++) A[pz] = E->argsN[pz];
    for (int pz = E->kpar; pz < kPar; pz++) A[pz] = E->args;
    for (int pz = 0; pz < kPar; pz++) {
        A[pz].in = static_cast<const uint8_t*>(d_frames);
        A[pz].arena = R.arena; A[pz].slice_size = R.slice_size; A[pz].slice_off = R.slice_off; A[pz].frame_off = R.frame_off;
        A[pz].frame_len = R.frame_len; A[pz].flags = R.flags;
    }
    CU(cudaMemsetAsync(A[0].flags, 0, 256, s));
    CU(cudaMemsetAsync(A[0].work_ctr, 0, (size_t)A[0].nbands * 4, s));
    // k_emit accumulates into the slice scratch: it is zeroed once at open and k_pack re-zeroes what a call has used; only a
    // call that ended in an overflow leaves it dirty
    if (E->scratch_dirty) {
        CU(cudaMemsetAsync(A[0].scratch, 0, (size_t)E->max_frames * A[0].nslices * A[0].slice_cap, s));
        E->scratch_dirty = false;
    }
    // Three kernels per band on three streams: model(b) -> range(b) -> emit(b); model(b) reuses the band buffers of
    // parity b&1 once emit(b-2) has drained them. `s` (the caller's stream) forks into and joins from the three.
    // timing mode (b200_ffv1_set_timing) and B200_SERIAL run everything on the caller's stream so that CUDA events can
    // bracket each kernel
    const bool serial = E->timing || getenv("B200_SERIAL") != nullptr;
    if (E->timing && E->tev.empty()) {
        E->tev.resize((size_t)A[0].nbands * 4 + 2);
        for (auto& ev : E->tev) CU(cudaEventCreate(&ev));
    }
    const bool tm = E->timing;
    const bool tr = !serial && getenv("B200_TRACE") != nullptr;
    if (tr && E->trace.empty()) {
        E->trace.resize((size_t)A[0].nbands * 6 + 1);
        for (auto& ev : E->trace) CU(cudaEventCreate(&ev));
    }
    if (tr) CU(cudaEventRecord(E->trace[(size_t)A[0].nbands * 6], s));
    E->trace_pending = tr;
    const int emode = serial ? 2 : E->emit_mode;     // serial: everything on the caller's stream, emit right behind range
    cudaStream_t sm = serial ? s : E->sm, sr = serial ? s : E->sr;
    cudaStream_t se = serial ? s : emode == 1 ? E->sm : emode == 2 ? E->sr : E->se;
    uint64_t launches = 0;
    if (!serial) {
        CU(cudaEventRecord(E->ev_start, s));
        CU(cudaStreamWaitEvent(sm, E->ev_start, 0));
        CU(cudaStreamWaitEvent(sr, E->ev_start, 0));
        if (emode == 0) CU(cudaStreamWaitEvent(se, E->ev_start, 0));
    }
    const int nb = A[0].nbands;
    std::vector<std::pair<int, int>> srows;                      // distinct (y0, h) of the slice grid's rows
    if (host_frames) {
        for (const auto& g : E->st.slices)
            if (srows.empty() || srows.back().first != g.y0) srows.push_back({g.y0, g.h});
        if (serial) prefetch = false;
        if (prefetch) {
            if (!(E->prefetched_set == set && E->prefetched_first == host_frames[0] && E->prefetched_n == n_frames)) {
                int r = prefetch_copies(E, (uint8_t*)d_frames, host_frames, n_frames, set);
                if (r) return r;
            }
            E->prefetched_set = -1;
            CU(cudaStreamWaitEvent(sm, E->ev_prefetch, 0));
        } else if (!serial) { CU(cudaEventRecord(E->ev_start, s)); CU(cudaStreamWaitEvent(E->sc, E->ev_start, 0)); }
    }
    // emit(band): after range(band); frees the band buffers of its set for model(band + kpar)
    auto do_emit = [&](int band) -> int {
        const int p = band % E->kpar;
        if (!serial && emode != 2) CU(cudaStreamWaitEvent(se, E->ev_range[p], 0));
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 4], se));
        if (!getenv("B200_SKIP_EMIT")) CU(b200::launch_emit(A[p], n_frames, se));     // (experiment knob: timing without the emitter; output invalid)
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 5], se));
        if (tm) CU(cudaEventRecord(E->tev[band * 4 + 3], s));
        if (!serial) CU(cudaEventRecord(E->ev_emit[p], se));
        launches++;
        return 0;
    };
    for (int band = 0; band < nb; band++) {
        const int p = band % E->kpar;
        if (host_frames && !prefetch) {
            cudaStream_t sc = serial ? s : E->sc;
            const size_t rb = E->st.row_bytes;
            for (int i = 0; i < n_frames; i++)
                for (const auto& sr_ : srows) {
                    const int r0 = band * A[0].band_rows;
                    if (r0 >= sr_.second) continue;
                    const int r1 = r0 + A[0].band_rows < sr_.second ? r0 + A[0].band_rows : sr_.second;
                    const size_t o = (size_t)(sr_.first + r0) * rb;
                    CU(cudaMemcpyAsync((uint8_t*)d_frames + (size_t)i * E->st.frame_bytes + o, host_frames[i] + o, (size_t)(r1 - r0) * rb,
                                       cudaMemcpyHostToDevice, sc));
                }
            if (!serial) { CU(cudaEventRecord(E->ev_h2d[band], sc)); CU(cudaStreamWaitEvent(sm, E->ev_h2d[band], 0)); }
        }
        // the band buffers of set p are free once emit(band - kpar) has drained them (emode 1: that emit is earlier in this stream)
        if (!serial && emode != 1 && band >= E->kpar) CU(cudaStreamWaitEvent(sm, E->ev_emit[p], 0));
        if (tm) CU(cudaEventRecord(E->tev[band * 4 + 0], s));
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 0], sm));
        CU(b200::launch_model(A[p], band, n_frames, sm));
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 1], sm));
        if (tm) CU(cudaEventRecord(E->tev[band * 4 + 1], s));
        if (!serial) { CU(cudaEventRecord(E->ev_model[p], sm)); CU(cudaStreamWaitEvent(sr, E->ev_model[p], 0)); }
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 2], sr));
        CU(b200::launch_range(A[p], band, n_frames, sr));
        if (tr) CU(cudaEventRecord(E->trace[band * 6 + 3], sr));
        if (tm) CU(cudaEventRecord(E->tev[band * 4 + 2], s));
        if (!serial) CU(cudaEventRecord(E->ev_range[p], sr));
        launches += 2;
        // emode 1: k_emit shares the model stream; emit(band - kpar + 1) goes in behind model(band), just ahead of the
        // model launch that needs its buffers, so that range(band - kpar + 1) has had a whole model launch to finish
        if (emode == 1) {
            const int eb = band - (E->kpar - 1);
            if (eb >= 0) { int r = do_emit(eb); if (r) return r; }
        } else {
            int r = do_emit(band); if (r) return r;
        }
    }
    if (emode == 1)
        for (int eb = nb - (E->kpar - 1) < 0 ? 0 : nb - (E->kpar - 1); eb < nb; eb++) { int r = do_emit(eb); if (r) return r; }
    if (!serial) {
        if (emode == 0) { CU(cudaEventRecord(E->ev_done_e, se)); CU(cudaStreamWaitEvent(s, E->ev_done_e, 0)); }
        CU(cudaEventRecord(E->ev_done_e, sr));
        CU(cudaStreamWaitEvent(s, E->ev_done_e, 0));
        CU(cudaEventRecord(E->ev_done_m, sm));
        CU(cudaStreamWaitEvent(s, E->ev_done_m, 0));
    }
    if (tm) CU(cudaEventRecord(E->tev[(size_t)nb * 4], s));
    CU(b200::launch_pack(A[0], n_frames, s));
    if (tm) CU(cudaEventRecord(E->tev[(size_t)nb * 4 + 1], s));
    E->timed_pending = tm;
    launches += 2;
    R.n_frames = n_frames;
    CU(cudaEventRecord(R.done, s));
    E->stats[0] = launches;
    E->stats[2] = (uint64_t)n_frames * E->st.width * E->st.height * 3;
    return 0;
}

int b200_ffv1_encode_device(b200_ffv1_enc* E, const void* d_frames, int32_t n_frames, void* stream) {
    if (!E || !d_frames) return fail(B200_ERR_INVALID, "null argument");
    if (n_frames < 1 || n_frames > E->max_frames) return fail(B200_ERR_INVALID, "n_frames out of range");
    E->fifo[0] = 0; E->fifo_n = 1; E->host_mode = false;   // the device entry point has one result set: the last call's
    return encode_impl(E, d_frames, n_frames, static_cast<cudaStream_t>(stream), nullptr, 0);
}

// waits for the oldest batch in flight and reads its tables; the batch stays at the head of the queue until `pop`
static int collect(b200_ffv1_enc* E, int32_t n_frames, size_t* out_off, size_t* out_len, uint64_t* total, bool pop, int* which) {
    if (E->fifo_n == 0) return fail(B200_ERR_INVALID, "no encode call to collect");
    const int set = E->fifo[0];
    b200_ffv1_enc::ResultSet& R = E->rs[set];
    if (which) *which = set;
    if (n_frames != R.n_frames) return fail(B200_ERR_INVALID, "n_frames differs from the encode call being collected");
    if (pop && E->host_mode) { E->fifo[0] = E->fifo[1]; E->fifo_n--; }
    const b200::EncArgs& A = E->args;
    CU(cudaEventSynchronize(R.done));
    CU(cudaMemcpyAsync(R.h_flags, R.flags, 256, cudaMemcpyDeviceToHost, E->sf));
    CU(cudaMemcpyAsync(R.h_off.data(), R.frame_off, (size_t)n_frames * 8, cudaMemcpyDeviceToHost, E->sf));
    CU(cudaMemcpyAsync(R.h_len.data(), R.frame_len, (size_t)n_frames * 8, cudaMemcpyDeviceToHost, E->sf));
    CU(cudaStreamSynchronize(E->sf));
    if (getenv("B200_PHASE_TIMING")) {
        const unsigned long long* ph = reinterpret_cast<const unsigned long long*>(R.h_flags + 16);
        fprintf(stderr, "k_model phase cycles (sum over CTAs): S1 %llu S2p %llu S2a %llu S2b %llu S3 %llu\n", ph[0], ph[1], ph[2], ph[3], ph[4]);
    }
    if (E->trace_pending) {
        const int nb = A.nbands;
        const cudaEvent_t t0 = E->trace[(size_t)nb * 6];
        fprintf(stderr, "band: model start-end | range start-end | emit start-end (ms)\n");
        for (int b = 0; b < nb; b++) {
            float v[6];
            for (int k = 0; k < 6; k++) { cudaEventSynchronize(E->trace[b * 6 + k]); cudaEventElapsedTime(&v[k], t0, E->trace[b * 6 + k]); }
            fprintf(stderr, "%2d: %7.2f-%7.2f | %7.2f-%7.2f | %7.2f-%7.2f\n", b, v[0], v[1], v[2], v[3], v[4], v[5]);
        }
        E->trace_pending = false;
    }
    if (R.h_flags[0] & 7u) E->scratch_dirty = true;
    if (R.h_flags[0] & 1u) return fail(B200_ERR_OVERFLOW, "slice scratch overflow");
    if (R.h_flags[0] & 2u) return fail(B200_ERR_OVERFLOW, "packet arena overflow");
    if (R.h_flags[0] & 4u) return fail(B200_ERR_OVERFLOW, "plane-row needs more column segments than reserved");
    uint64_t tot = 0;
    for (int i = 0; i < n_frames; i++) {
        if (out_off) out_off[i] = (size_t)R.h_off[i];
        if (out_len) out_len[i] = (size_t)R.h_len[i];
        tot += R.h_len[i];
    }
    E->stats[1] = (uint64_t)R.h_flags[2] | ((uint64_t)R.h_flags[3] << 32);
    if (E->timed_pending) {      // device time of each kernel class in the last (serial, timed) encode, microseconds
        double tmod = 0, trng = 0, temt = 0, tpk = 0;
        float ms = 0;
        const int nb = A.nbands;
        CU(cudaEventSynchronize(E->tev[(size_t)nb * 4 + 1]));
        for (int b = 0; b < nb; b++) {
            cudaEventElapsedTime(&ms, E->tev[b * 4 + 0], E->tev[b * 4 + 1]); tmod += ms;
            cudaEventElapsedTime(&ms, E->tev[b * 4 + 1], E->tev[b * 4 + 2]); trng += ms;
            cudaEventElapsedTime(&ms, E->tev[b * 4 + 2], E->tev[b * 4 + 3]); temt += ms;
        }
        cudaEventElapsedTime(&ms, E->tev[(size_t)nb * 4], E->tev[(size_t)nb * 4 + 1]); tpk = ms;
        E->stats[4] = (uint64_t)(tmod * 1000); E->stats[5] = (uint64_t)(trng * 1000); E->stats[6] = (uint64_t)(tpk * 1000);
        E->stats[7] = (uint64_t)(temt * 1000);
        E->timed_pending = false;
    }
    E->stats[3] = tot;
    if (total) *total = tot;
    return 0;
}

int b200_ffv1_packets_device(b200_ffv1_enc* E, const void** d_arena, size_t* out_off, size_t* out_len, int32_t n_frames) {
    if (!E) return fail(B200_ERR_INVALID, "null encoder");
    CU(cudaSetDevice(E->cfg.device));
    // the batch leaves the queue: the caller reads its arena before submitting again (with two batches in flight the arena
    // returned here is the one the NEXT submit will not touch)
    int set = 0;
    int r = collect(E, n_frames, out_off, out_len, nullptr, true, &set);
    if (r) return r;
    if (d_arena) *d_arena = E->rs[set].arena;
    return 0;
}

int b200_ffv1_fetch_packets(b200_ffv1_enc* E, uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len, int32_t n_frames) {
    if (!E || !out) return fail(B200_ERR_INVALID, "null argument");
    CU(cudaSetDevice(E->cfg.device));
    uint64_t total = 0;
    int set = 0;
    int r = collect(E, n_frames, out_off, out_len, &total, true, &set);
    if (r) return r;
    if (total > out_cap) return fail(B200_ERR_OVERFLOW, "output buffer too small");
    CU(cudaMemcpyAsync(out, E->rs[set].arena, (size_t)total, cudaMemcpyDeviceToHost, E->sf));
    CU(cudaStreamSynchronize(E->sf));
    return 0;
}

int b200_ffv1_packet_sizes(b200_ffv1_enc* E, size_t* out_off, size_t* out_len, int32_t n_frames, size_t* total_bytes) {
    if (!E) return fail(B200_ERR_INVALID, "null encoder");
    CU(cudaSetDevice(E->cfg.device));
    uint64_t total = 0;
    int r = collect(E, n_frames, out_off, out_len, &total, false, nullptr);
    if (r) return r;
    if (total_bytes) *total_bytes = (size_t)total;
    return 0;
}

int b200_ffv1_prefetch_host(b200_ffv1_enc* E, const uint8_t* const* frames, int32_t n_frames) {
    if (!E || !frames) return fail(B200_ERR_INVALID, "null argument");
    if (n_frames < 1 || n_frames > E->max_frames) return fail(B200_ERR_INVALID, "n_frames out of range");
    CU(cudaSetDevice(E->cfg.device));
    for (int i = 0; i < n_frames; i++)
        if (!frames[i]) return fail(B200_ERR_INVALID, "null frame pointer");
    if (!E->host_mode || E->fifo_n == 0 || E->timing) return 0;         // nothing in flight: the submit copies band by band
    // the staging buffer (= result set) the next submit will take: the one the newest batch in flight does not use
    const int set = E->fifo[E->fifo_n - 1] ^ 1;
    uint8_t** buf = set ? &E->d_in2 : &E->d_in;
    if (!*buf) CU(cudaMalloc((void**)buf, E->st.frame_bytes * E->max_frames));
    int r = prefetch_copies(E, *buf, frames, n_frames, set);
    if (r) return r;
    E->prefetched_set = set; E->prefetched_first = frames[0]; E->prefetched_n = n_frames;
    return 0;
}

int b200_ffv1_submit_host(b200_ffv1_enc* E, const uint8_t* const* frames, int32_t n_frames) {
    if (!E || !frames) return fail(B200_ERR_INVALID, "null argument");
    if (n_frames < 1 || n_frames > E->max_frames) return fail(B200_ERR_INVALID, "n_frames out of range");
    CU(cudaSetDevice(E->cfg.device));
    const size_t fb = E->st.frame_bytes;
    for (int i = 0; i < n_frames; i++)
        if (!frames[i]) return fail(B200_ERR_INVALID, "null frame pointer");
    // result set: 0 when nothing is in flight; the free one when one batch is; with two in flight the oldest is given up
    if (!E->host_mode) { E->fifo_n = 0; E->host_mode = true; }
    int set = 0;
    if (E->fifo_n == 1) set = E->fifo[0] ^ 1;
    else if (E->fifo_n == 2) { set = E->fifo[0]; E->fifo[0] = E->fifo[1]; E->fifo_n = 1; }
    // one staging buffer per result set; with a batch in flight the new one is prefetched into the other buffer
    const bool prefetch = E->fifo_n >= 1 && !E->timing && !getenv("B200_NO_PREFETCH");
    E->fifo[E->fifo_n++] = set;
    uint8_t** buf = set ? &E->d_in2 : &E->d_in;
    if (!*buf) CU(cudaMalloc((void**)buf, fb * E->max_frames));
    return encode_impl(E, *buf, n_frames, E->sh, frames, set, prefetch);
}

int b200_ffv1_encode_host(b200_ffv1_enc* E, const uint8_t* const* frames, int32_t n_frames,
                          uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len) {
    if (!out) return fail(B200_ERR_INVALID, "null argument");
    int r = b200_ffv1_submit_host(E, frames, n_frames);
    if (r) return r;
    return b200_ffv1_fetch_packets(E, out, out_cap, out_off, out_len, n_frames);
}

int b200_ffv1_info(const b200_ffv1_enc* E, int32_t info[8]) {
    if (!E || !info) return fail(B200_ERR_INVALID, "null argument");
    info[0] = E->st.num_h; info[1] = E->st.num_v; info[2] = E->args.nbands; info[3] = E->args.band_rows;
    info[4] = E->args.nslices; info[5] = E->args.wmax; info[6] = E->args.hmax; info[7] = E->st.bits;
    return 0;
}

int b200_ffv1_stats(const b200_ffv1_enc* E, uint64_t stats[8]) {
    if (!E || !stats) return fail(B200_ERR_INVALID, "null argument");
    std::memcpy(stats, E->stats, sizeof E->stats);
    return 0;
}

}  // extern "C"
