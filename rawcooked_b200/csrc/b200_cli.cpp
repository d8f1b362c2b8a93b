// b200enc_main: the `ffmpeg`-compatible entry point RAWcooked launches (`rawcooked -b b200enc ...`).
//
// Understands exactly the argv grammar /root/reference/Source/CLI/Output.cpp:36-378 emits:
//   <bin> -xerror [-nostdin] { [-framerate F -r F] -f image2|concat [-safe 0] -c:v dpx|tiff [-start_number N] -i <pattern|list> | -i <wav> }+
//         [-map K]... -c:a flac|copy -c:v ffv1 -coder 1 -context C -f matroska -g 1 -level 3 [-loglevel L] [-n|-y] -slicecrc S
//         [-slices N] [-threads T] { -attach <file> -metadata:s:K mimetype=... -metadata:s:K filename=<name> }*  -f matroska <out.mkv>
// re-opens the source files itself, encodes video on the GPU through the C ABI and writes the Matroska file with the
// attachments ahead of the first Cluster. Exit status: 0 on success, non-zero otherwise (Output.cpp:356-360).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200dec.h"
#include "../../include/b200enc.h"
#include "../../include/b200scan.h"
#include "ingest.h"
#include "mkv_mux.h"

namespace {

struct InputSpec {
    std::string path;
    std::map<std::string, std::string> opt;   // options that preceded this -i
};
struct AttachSpec {
    std::string path, name, mime;
};

bool read_at(int fd, void* dst, size_t n, uint64_t off) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
        ssize_t r = pread(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; n -= (size_t)r; off += (uint64_t)r;
    }
    return true;
}
bool read_file(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    bool ok = n <= 0 || fread(out->data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}
uint64_t file_size(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0; }
double parse_rate(const std::string& s) {
    const size_t sl = s.find('/');
    if (sl == std::string::npos) return atof(s.c_str());
    const double d = atof(s.c_str() + sl + 1);
    return d ? atof(s.substr(0, sl).c_str()) / d : 0;
}

// "24", "24000/1001" or "23.976" as a reduced fraction (the time base of a framemd5 output is its inverse)
void parse_rate_q(const std::string& s, long long* num, long long* den) {
    long long n = 0, d = 1;
    const size_t sl = s.find('/');
    if (sl != std::string::npos) { n = atoll(s.c_str()); d = atoll(s.c_str() + sl + 1); }
    else if (s.find('.') != std::string::npos) { n = llround(atof(s.c_str()) * 1000000.0); d = 1000000; }
    else n = atoll(s.c_str());
    if (n <= 0 || d <= 0) { n = 25; d = 1; }
    long long a = n, b = d;
    while (b) { const long long t = a % b; a = b; b = t; }
    *num = n / a; *den = d / a;
}

struct VideoStream {
    std::vector<std::string> files;
    b200::ImageInfo info;
    double fps = 24;
    long long fps_num = 24, fps_den = 1;
    char kind = 'd';
    int track = 0;
};
struct AudioStream {
    std::string path;
    b200::WavInfo info;
    int track = 0;
    bool flac = true;
};

// B200_CLI_TIMING=1: wall-clock of the front-end's phases on stderr
struct PhaseClock {
    bool on = getenv("B200_CLI_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "b200enc: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

int fail(const std::string& msg, int code = 1) {
    fprintf(stderr, "b200enc: %s\n", msg.c_str());
    return code;
}

}  // namespace


// set by the b200enc executable: after the Matroska file is complete the process exits without freeing device memory and
// unpinning host memory (seconds for tens of GB); a library caller of b200enc_main gets the full teardown
bool g_fast_exit = false;
extern "C" void b200enc_set_fast_exit(int on) { g_fast_exit = on != 0; }

namespace {

// One worker per GPU. Batch k (frames [k*B, k*B + B)) belongs to worker k % G, which keeps two of its batches in flight
// through b200_ffv1_submit_host: while batch j is being coded, the files of batch j+1 are read into the other pinned input
// buffer and the packets of batch j-1 cross PCIe into a pinned output buffer. The muxer (the calling thread) takes the
// batches in order and hands the packets to the Matroska writer's pwrite pool straight from those buffers.
struct Batch {
    size_t f0 = 0, n = 0;
    uint8_t* data = nullptr;                  // pinned output buffer of the owning worker
    std::vector<size_t> off, len;
    bool ready = false, consumed = false;
    std::string err;
};

struct VideoPipeline {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Batch> batches;
    bool abort = false;
};

struct PinnedBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t need) {
        if (need <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = need + (need >> 3);
        if (cudaHostAlloc((void**)&p, want, cudaHostAllocDefault) != cudaSuccess) return false;
        cap = want;
        return true;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

template <class FlushAudio>
int encode_video_stream(VideoStream& v, b200_ffv1_enc* first_enc, const std::vector<int>& devices, int frames_in_flight, unsigned nthreads,
                        int slices, int context, int slicecrc, b200::MkvWriter& mux, FlushAudio& flush_audio_until, PhaseClock& pc, std::string* err,
                        std::vector<uint8_t>* frame_md5 /* null, or 16 bytes per frame: the digests of a framemd5 output */) {
    const size_t fb = b200_ffv1_frame_bytes(v.info.width, v.info.height, v.info.layout);
    const size_t N = v.files.size();
    const size_t B = std::min<size_t>((size_t)frames_in_flight, N);
    const char* ve = getenv("B200_VERIFY");
    const bool verify = ve && atoi(ve) > 0;
    std::atomic<size_t> verified{0};
    const size_t nbatch = (N + B - 1) / B;
    const size_t G = std::max<size_t>(1, std::min(devices.size(), nbatch));
    VideoPipeline P;
    P.batches.resize(nbatch);
    for (size_t k = 0; k < nbatch; k++) { P.batches[k].f0 = k * B; P.batches[k].n = std::min(B, N - k * B); }
    const unsigned nread = std::max(1u, std::min(nthreads / (unsigned)G, 32u));

    // reads frames [f0, f0 + n) into dst (pread straight into pinned memory; one header parse per file because the payload
    // offset may differ from frame to frame); returns "" or the first error
    auto read_batch = [&](size_t f0, size_t n, uint8_t* dst) -> std::string {
        std::vector<std::string> errs(nread);
        auto work = [&](unsigned t) {
            std::vector<uint8_t> head(65536);
            for (size_t k = t; k < n && errs[t].empty(); k += nread) {
                const std::string& path = v.files[f0 + k];
                const int fd = open(path.c_str(), O_RDONLY);
                if (fd < 0) { errs[t] = "cannot open " + path; return; }
                b200::ImageInfo fi = v.info;
                if (f0 + k > 0) {
                    const ssize_t hn = pread(fd, head.data(), head.size(), 0);
                    std::string e;
                    const bool ok = v.kind == 'd' ? b200::parse_dpx(head.data(), hn > 0 ? (size_t)hn : 0, file_size(path), &fi, &e)
                                                  : b200::parse_tiff(head.data(), hn > 0 ? (size_t)hn : 0, file_size(path), &fi, &e);
                    if (!ok || fi.width != v.info.width || fi.height != v.info.height || fi.layout != v.info.layout) {
                        close(fd);
                        errs[t] = path + ": " + (ok ? std::string("geometry differs from the first frame") : e);
                        return;
                    }
                }
                if (!read_at(fd, dst + k * fb, fb, fi.data_offset)) { close(fd); errs[t] = "cannot read " + path; return; }
                close(fd);
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nread; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (auto& e : errs) if (!e.empty()) return e;
        return "";
    };

    auto worker = [&](size_t w, b200_ffv1_enc* E) {
        const int device = devices[w];
        std::string werr;
        auto fail_all = [&](const std::string& m) {
            std::lock_guard<std::mutex> lk(P.mu);
            for (size_t k = w; k < nbatch; k += G) if (!P.batches[k].ready) { P.batches[k].err = m; P.batches[k].ready = true; }
            P.abort = true;
            P.cv.notify_all();
        };
        cudaSetDevice(device);
        if (!E) {
            b200_ffv1_cfg cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.width = v.info.width; cfg.height = v.info.height; cfg.layout = v.info.layout;
            cfg.slices = slices; cfg.context = context; cfg.coder = 1; cfg.slicecrc = slicecrc;
            cfg.max_frames = (int32_t)B; cfg.device = device;
            if (b200_ffv1_open(&cfg, &E)) { fail_all(std::string("ffv1: ") + b200_last_error()); return; }
        }
        // B200_VERIFY=1: every packet is decoded again on the GPU (k_decode, the `--check` decoder of include/b200dec.h) and
        // compared with the source payload it was coded from, while the next batch is being coded; any difference fails the job
        b200_ffv1_dec* D = nullptr;
        if (verify) {
            std::vector<uint8_t> rec(b200_ffv1_config_record(E, nullptr, 0));
            b200_ffv1_config_record(E, rec.data(), rec.size());
            b200_ffv1_dec_cfg dc;
            memset(&dc, 0, sizeof dc);
            dc.width = v.info.width; dc.height = v.info.height; dc.layout = v.info.layout; dc.max_frames = (int32_t)B; dc.device = device;
            if (b200_ffv1_dec_open(&dc, rec.data(), rec.size(), &D)) { b200_ffv1_close(E); fail_all(std::string("verify: ") + b200_last_error()); return; }
        }
        struct DecGuard { b200_ffv1_dec*& d; ~DecGuard() { if (d && !g_fast_exit) b200_ffv1_dec_close(d); } } dec_guard{D};
        // `-f framemd5`: MD5 of every frame in FFmpeg's rawvideo form (k_rawframe + k_md5, include/b200scan.h), one lane per frame
        b200_scan* Sc = nullptr;
        if (frame_md5 && b200_scan_open(device, (int32_t)B, fb * B, &Sc)) { b200_ffv1_close(E); fail_all(std::string("framemd5: ") + b200_last_error()); return; }
        struct ScanGuard { b200_scan*& s; ~ScanGuard() { if (s && !g_fast_exit) b200_scan_close(s); } } scan_guard{Sc};
        PinnedBuf in[2], out[2];
        // pinning is slow (a few GB/s): only the first input buffer is pinned before work starts; the second one and the
        // output buffers are pinned by background threads while the first batch is read and coded, each joined where its
        // buffer is first needed
        const size_t out_guess = fb * B - (fb * B >> 2);
        const bool many = (nbatch + G - 1 - w) / G > 1;
        bool bg_ok[3] = {true, true, true};
        std::thread bg[3];
        struct JoinGuard { std::thread* t; ~JoinGuard() { for (int i = 0; i < 3; i++) if (t[i].joinable()) t[i].join(); } } guard{bg};
        auto join_bg = [&](int i) -> bool { if (bg[i].joinable()) bg[i].join(); return bg_ok[i]; };
        if (!in[0].ensure(fb * B)) { b200_ffv1_close(E); fail_all("cannot allocate pinned host buffers"); return; }
        bg[1] = std::thread([&] { cudaSetDevice(device); bg_ok[1] = out[0].ensure(out_guess); });
        if (many) {
            bg[0] = std::thread([&] { cudaSetDevice(device); bg_ok[0] = in[1].ensure(fb * B); });
            bg[2] = std::thread([&] { cudaSetDevice(device); bg_ok[2] = out[1].ensure(out_guess); });
        }
        std::vector<const uint8_t*> ptrs(B);
        size_t j = 0;
        long prev = -1;                                           // my previous batch: submitted, not yet fetched
        auto fetch = [&](size_t k, size_t jj) -> bool {
            Batch& bt = P.batches[k];
            bt.off.resize(bt.n); bt.len.resize(bt.n);
            size_t total = 0;
            if (b200_ffv1_packet_sizes(E, bt.off.data(), bt.len.data(), (int32_t)bt.n, &total)) { werr = std::string("ffv1 encode: ") + b200_last_error(); return false; }
            PinnedBuf& o = out[jj & 1];
            if (jj >= 2) {                                        // the muxer must be done with the batch that used this buffer
                std::unique_lock<std::mutex> lk(P.mu);
                P.cv.wait(lk, [&] { return P.batches[k - 2 * G].consumed || P.abort; });
                if (P.abort) { werr = "aborted"; return false; }
            }
            if (!join_bg(1 + (int)(jj & 1)) || !o.ensure(total)) { werr = "cannot allocate pinned host buffers"; return false; }
            if (b200_ffv1_fetch_packets(E, o.p, o.cap, bt.off.data(), bt.len.data(), (int32_t)bt.n)) { werr = std::string("ffv1 fetch: ") + b200_last_error(); return false; }
            if (Sc) {
                std::vector<const uint8_t*> src(bt.n);
                for (size_t i = 0; i < bt.n; i++) src[i] = in[jj & 1].p + i * fb;
                if (b200_framemd5_host(Sc, v.info.width, v.info.height, v.info.layout, src.data(), (int32_t)bt.n, frame_md5->data() + 16 * bt.f0)) {
                    werr = std::string("framemd5: ") + b200_last_error();
                    return false;
                }
            }
            if (D) {
                std::vector<const uint8_t*> pk(bt.n), src(bt.n);
                std::vector<uint64_t> mm(bt.n);
                std::vector<uint32_t> st(bt.n);
                for (size_t i = 0; i < bt.n; i++) { pk[i] = o.p + bt.off[i]; src[i] = in[jj & 1].p + i * fb; }
                if (b200_ffv1_check_host(D, pk.data(), bt.len.data(), (int32_t)bt.n, src.data(), mm.data(), st.data())) {
                    werr = std::string("verify: ") + b200_last_error();
                    return false;
                }
                for (size_t i = 0; i < bt.n; i++)
                    if (mm[i] || st[i]) {
                        werr = "verify: " + v.files[bt.f0 + i] + " does not decode to itself (" + std::to_string(mm[i]) + " bytes differ, decoder status " + std::to_string(st[i]) + ")";
                        return false;
                    }
                verified += bt.n;
            }
            {
                std::lock_guard<std::mutex> lk(P.mu);
                bt.data = o.p;
                bt.ready = true;
            }
            P.cv.notify_all();
            return true;
        };
        for (size_t k = w; k < nbatch; k += G, j++) {
            Batch& bt = P.batches[k];
            if (j == 1 && !join_bg(0)) { werr = "cannot allocate pinned host buffers"; break; }
            const std::string rerr = read_batch(bt.f0, bt.n, in[j & 1].p);
            if (!rerr.empty()) { werr = rerr; break; }
            for (size_t i = 0; i < bt.n; i++) ptrs[i] = in[j & 1].p + i * fb;
            if (b200_ffv1_submit_host(E, ptrs.data(), (int32_t)bt.n)) { werr = std::string("ffv1 encode: ") + b200_last_error(); break; }
            if (prev >= 0 && !fetch((size_t)prev, j - 1)) break;
            prev = (long)k;
            {
                std::lock_guard<std::mutex> lk(P.mu);
                if (P.abort) { werr = "aborted"; break; }
            }
        }
        if (werr.empty() && prev >= 0) fetch((size_t)prev, j - 1);
        if (!werr.empty()) fail_all(werr);
        // the pinned output buffers must outlive the muxer's use of them
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] {
                if (P.abort) return true;
                for (size_t k = w; k < nbatch; k += G) if (!P.batches[k].consumed) return false;
                return true;
            });
        }
        for (int i = 0; i < 3; i++) join_bg(i);
        if (g_fast_exit) { in[0].p = in[1].p = out[0].p = out[1].p = nullptr; return; }   // the process is about to end: no teardown
        b200_ffv1_close(E);
    };

    std::vector<std::thread> workers;
    for (size_t w = 0; w < G; w++) workers.emplace_back(worker, w, w == 0 ? first_enc : nullptr);
    pc.mark("workers started");
    int rc = 0;
    for (size_t k = 0; k < nbatch && !rc; k++) {
        Batch& bt = P.batches[k];
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] { return bt.ready; });
        }
        pc.mark("wait for batch");
        if (!bt.err.empty()) { *err = bt.err; rc = 1; break; }
        bool ok = true;
        for (size_t i = 0; i < bt.n && ok; i++) {
            const int64_t t_ms = (int64_t)std::llround(1000.0 * (double)(bt.f0 + i) / v.fps);
            ok = flush_audio_until(t_ms) && mux.write_block(v.track, t_ms, bt.data + bt.off[i], bt.len[i]);
        }
        ok = ok && mux.sync();                                    // the packets have left the pinned buffer
        pc.mark("mux write (batch)");
        if (!ok) { *err = mux.error(); rc = B200_ERR_IO; }
        {
            std::lock_guard<std::mutex> lk(P.mu);
            bt.consumed = true;
            if (rc) P.abort = true;
        }
        P.cv.notify_all();
    }
    if (rc) {
        std::lock_guard<std::mutex> lk(P.mu);
        P.abort = true;
        P.cv.notify_all();
    }
    for (auto& t : workers) t.join();
    if (!rc && verify)
        fprintf(stderr, "b200enc: %zu of %zu frames decoded again on the GPU and compared with their source payloads: no difference\n", verified.load(), N);
    return rc;
}

}  // namespace

int b200_flac_encode_file_to_mux(const std::string& path, const b200::WavInfo& wi, int track, int device,
                                 std::vector<uint8_t>* codec_private, std::vector<std::pair<int64_t, std::vector<uint8_t>>>* packets,
                                 std::string* err);   // flac_host.cpp

extern "C" int b200enc_main(int argc, char** argv) {
    PhaseClock pc;
    std::vector<InputSpec> inputs;
    std::vector<AttachSpec> attaches;
    std::map<std::string, std::string> pending, outopt;
    std::vector<std::string> outputs;
    std::string framemd5_path;
    bool overwrite = false, never = false;
    int cur_attach = -1;
    bool after_inputs = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-xerror" || a == "-nostdin" || a == "-an" || a == "-hide_banner") continue;
        if (a == "-y") { overwrite = true; continue; }
        if (a == "-n") { never = true; continue; }
        if (a.size() > 1 && a[0] == '-') {
            if (i + 1 >= argc) return fail("option " + a + " needs a value");
            const std::string v = argv[++i];
            if (a == "-i") { inputs.push_back({v, pending}); pending.clear(); continue; }
            if (a == "-attach") { attaches.push_back({v, "", ""}); cur_attach = (int)attaches.size() - 1; after_inputs = true; continue; }
            if (a.compare(0, 12, "-metadata:s:") == 0) {
                if (cur_attach >= 0) {
                    if (v.compare(0, 9, "filename=") == 0) attaches[cur_attach].name = v.substr(9);
                    else if (v.compare(0, 9, "mimetype=") == 0) attaches[cur_attach].mime = v.substr(9);
                }
                continue;   // -metadata:s:v WARNING=... and friends carry nothing the bitstream needs
            }
            if (a == "-map") { after_inputs = true; continue; }
            pending[a] = v;
            (void)after_inputs;
            continue;
        }
        // a bare token is an output file; the options gathered since the last -i / output belong to it. The first output is
        // the Matroska file; RAWcooked's --framemd5 adds `[-an] -f framemd5 <file>` as a second one (Output.cpp:312-332)
        if (outputs.empty()) {
            for (auto& kv : pending) outopt[kv.first] = kv.second;
        } else {
            const auto f = pending.find("-f");
            if (f == pending.end() || f->second != "framemd5" || !framemd5_path.empty())
                return fail("only one Matroska output and one `-f framemd5` output are supported");
            framemd5_path = a;
        }
        pending.clear();
        outputs.push_back(a);
    }
    if (inputs.empty()) return fail("no input (-i)");
    if (outputs.empty()) return fail("no output file");
    if (outopt.count("-f") && outopt["-f"] != "matroska") return fail("only -f matroska is supported");
    if (outopt.count("-c:v") && outopt["-c:v"] != "ffv1") return fail("only -c:v ffv1 is supported");
    if (outopt.count("-coder") && outopt["-coder"] != "1") return fail("only -coder 1 is supported");
    if (outopt.count("-level") && outopt["-level"] != "3") return fail("only -level 3 is supported");
    if (outopt.count("-g") && outopt["-g"] != "1") return fail("only -g 1 is supported");
    if (outopt.count("-vf")) return fail("-vf filters are not supported");
    const std::string out_path = outputs[0];
    if (file_size(out_path) > 0 || access(out_path.c_str(), F_OK) == 0) {
        if (never) return fail("File '" + out_path + "' already exists. Exiting.");
        if (!overwrite) return fail("File '" + out_path + "' already exists (use -y).");
    }
    const int context = outopt.count("-context") ? atoi(outopt["-context"].c_str()) : 0;
    const int slicecrc = outopt.count("-slicecrc") ? atoi(outopt["-slicecrc"].c_str()) : 1;
    const int slices = outopt.count("-slices") ? atoi(outopt["-slices"].c_str()) : 0;
    const bool audio_flac = !(outopt.count("-c:a") && outopt["-c:a"] == "copy");
    // GPUs: B200_DEVICES="0,1,..." (an ordinal may repeat: two workers on one GPU), else B200_DEVICE=k, else every visible device
    std::vector<int> devices;
    if (const char* e = getenv("B200_DEVICES")) {
        for (const char* q = e; *q;) {
            char* endp = nullptr;
            const long d = strtol(q, &endp, 10);
            if (endp == q) break;
            devices.push_back((int)d);
            q = *endp == ',' ? endp + 1 : endp;
        }
    } else if (const char* e1 = getenv("B200_DEVICE")) {
        devices.push_back(atoi(e1));
    } else {
        int nd = 0;
        if (cudaGetDeviceCount(&nd) == cudaSuccess) for (int d = 0; d < nd; d++) devices.push_back(d);
    }
    if (devices.empty()) devices.push_back(0);
    const int device = devices[0];
    unsigned nthreads = std::thread::hardware_concurrency();
    if (outopt.count("-threads") && atoi(outopt["-threads"].c_str()) > 0) nthreads = (unsigned)atoi(outopt["-threads"].c_str());
    nthreads = std::max(1u, nthreads);

    // ---- classify the inputs
    std::vector<VideoStream> videos;
    std::vector<AudioStream> audios;
    std::vector<std::pair<char, int>> order;      // stream order = input order = Matroska track order
    for (const InputSpec& in : inputs) {
        std::vector<std::string> files;
        const auto f = in.opt.find("-f");
        if (f != in.opt.end() && f->second == "concat") files = b200::read_concat_list(in.path);
        else if (f != in.opt.end() && f->second == "image2") {
            const auto sn = in.opt.find("-start_number");
            files = b200::expand_image2(in.path, sn != in.opt.end() ? atoll(sn->second.c_str()) : 0);
        } else files.push_back(in.path);
        if (files.empty()) return fail("no file matches " + in.path);
        uint8_t head[65536];
        const int fd = open(files[0].c_str(), O_RDONLY);
        if (fd < 0) return fail("cannot open " + files[0], B200_ERR_IO);
        const ssize_t n = pread(fd, head, sizeof head, 0);
        close(fd);
        const char kind = b200::sniff(head, n > 0 ? (size_t)n : 0);
        std::string err;
        if (kind == 'd' || kind == 't') {
            VideoStream v;
            v.files = files; v.kind = kind;
            const bool ok = kind == 'd' ? b200::parse_dpx(head, (size_t)n, file_size(files[0]), &v.info, &err)
                                        : b200::parse_tiff(head, (size_t)n, file_size(files[0]), &v.info, &err);
            if (!ok) return fail(files[0] + ": " + err);
            const auto fr = in.opt.find("-framerate");
            const auto r = in.opt.find("-r");
            v.fps = fr != in.opt.end() ? parse_rate(fr->second) : r != in.opt.end() ? parse_rate(r->second) : 25.0;   // ffmpeg's image2 default
            if (v.fps <= 0) v.fps = 25.0;
            parse_rate_q(fr != in.opt.end() ? fr->second : r != in.opt.end() ? r->second : std::string("25"), &v.fps_num, &v.fps_den);
            order.push_back({'v', (int)videos.size()});
            videos.push_back(v);
        } else if (kind == 'w') {
            AudioStream a;
            a.path = files[0];
            if (!b200::parse_wav(head, (size_t)n, file_size(files[0]), &a.info, &err)) return fail(files[0] + ": " + err);
            a.flac = audio_flac;
            if (a.flac && (a.info.is_float || a.info.bits > 24)) return fail(files[0] + ": FLAC needs integer PCM of at most 24 bits (use -c:a copy)");
            order.push_back({'a', (int)audios.size()});
            audios.push_back(a);
        } else {
            return fail(files[0] + ": unsupported input format");
        }
    }

    // several video inputs (RAWcooked emits one per run of consecutive frame numbers when a sequence has gaps, gaps.sh) become
    // several V_FFV1 tracks, muxed one after the other: each track's timestamps start at 0, which the reference's parser and
    // --check accept (tests/test_cli_gpu.py::test_rawcooked_gaps_concat_lists) although players would prefer them interleaved

    // ---- attachments
    std::vector<b200::MkvAttachment> atts;
    for (const AttachSpec& a : attaches) {
        b200::MkvAttachment m;
        m.name = a.name.empty() ? a.path.substr(a.path.find_last_of('/') + 1) : a.name;
        m.mime = a.mime;
        if (!read_file(a.path, &m.data)) return fail("cannot read attachment " + a.path, B200_ERR_IO);
        atts.push_back(std::move(m));
    }

    // ---- tracks
    std::vector<b200::MkvTrack> tracks;
    std::vector<b200_ffv1_enc*> encs(videos.size(), nullptr);
    std::vector<std::vector<std::pair<int64_t, std::vector<uint8_t>>>> audio_packets(audios.size());
    double duration_ms = 0;
    // throughput comes from the batch (k_range's serial time per band is fixed) but every frame in flight costs
    // pinned host memory and device memory that take seconds to set up: 64 is the better trade below a few thousand frames
    int frames_in_flight = 64;
    if (const char* e = getenv("B200_FRAMES_IN_FLIGHT")) frames_in_flight = std::max(1, atoi(e));
    auto cleanup = [&]() { for (auto* e : encs) b200_ffv1_close(e); };
    for (auto& o : order) {
        b200::MkvTrack t;
        if (o.first == 'v') {
            VideoStream& v = videos[o.second];
            b200_ffv1_cfg cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.width = v.info.width; cfg.height = v.info.height; cfg.layout = v.info.layout;
            cfg.slices = slices; cfg.context = context; cfg.coder = 1; cfg.slicecrc = slicecrc;
            cfg.max_frames = (int32_t)std::min<size_t>((size_t)frames_in_flight, v.files.size());
            cfg.device = device;
            const int rc = b200_ffv1_open(&cfg, &encs[o.second]);
            if (rc) { cleanup(); return fail(std::string("ffv1: ") + b200_last_error(), rc == B200_ERR_NO_DEVICE ? 2 : 1); }
            t.video = true; t.codec_id = "V_FFV1";
            t.codec_private.resize(b200_ffv1_config_record(encs[o.second], nullptr, 0));
            b200_ffv1_config_record(encs[o.second], t.codec_private.data(), t.codec_private.size());
            t.width = v.info.width; t.height = v.info.height; t.frame_rate = v.fps;
            duration_ms = std::max(duration_ms, 1000.0 * v.files.size() / v.fps);
            v.track = (int)tracks.size() + 1;
        } else {
            AudioStream& a = audios[o.second];
            a.track = (int)tracks.size() + 1;
            t.video = false;
            t.sample_rate = a.info.sample_rate; t.channels = a.info.channels; t.bit_depth = a.info.bits;
            const uint64_t nsamp = a.info.data_bytes / (a.info.channels * (a.info.bits / 8));
            duration_ms = std::max(duration_ms, 1000.0 * nsamp / a.info.sample_rate);
            if (a.flac) {
                t.codec_id = "A_FLAC";
                std::string err;
                const int rc = b200_flac_encode_file_to_mux(a.path, a.info, a.track, device, &t.codec_private, &audio_packets[o.second], &err);
                if (rc) { cleanup(); return fail("flac: " + err, rc == B200_ERR_NO_DEVICE ? 2 : 1); }
            } else {
                t.codec_id = a.info.is_float ? "A_PCM/FLOAT/IEEE" : "A_PCM/INT/LIT";
                // PCM copy: packets of ~1/25 s straight from the data chunk
                std::vector<uint8_t> raw(a.info.data_bytes);
                const int fd = open(a.path.c_str(), O_RDONLY);
                if (fd < 0 || !read_at(fd, raw.data(), raw.size(), a.info.data_offset)) { if (fd >= 0) close(fd); cleanup(); return fail("cannot read " + a.path, B200_ERR_IO); }
                close(fd);
                const size_t bpf = a.info.channels * (a.info.bits / 8);
                const size_t per = std::max<size_t>(1, a.info.sample_rate / 25);
                for (size_t s = 0; s * bpf < raw.size(); s += per) {
                    const size_t b0 = s * bpf, b1 = std::min(raw.size(), (s + per) * bpf);
                    audio_packets[o.second].push_back({(int64_t)std::llround(1000.0 * s / a.info.sample_rate), std::vector<uint8_t>(raw.begin() + b0, raw.begin() + b1)});
                }
            }
        }
        tracks.push_back(std::move(t));
    }

    pc.mark("parse + encoder open");
    b200::MkvWriter mux;
    if (!mux.open(out_path, tracks, atts, duration_ms)) { cleanup(); return fail(mux.error(), B200_ERR_IO); }

    // ---- encode: video in batches through the GPU(s), audio packets interleaved by timestamp
    std::vector<size_t> apos(audios.size(), 0);
    auto flush_audio_until = [&](int64_t t_ms) -> bool {
        for (size_t k = 0; k < audios.size(); k++)
            while (apos[k] < audio_packets[k].size() && audio_packets[k][apos[k]].first <= t_ms) {
                auto& p = audio_packets[k][apos[k]++];
                if (!mux.write_block(audios[k].track, p.first, p.second.data(), p.second.size())) return false;
            }
        return true;
    };
    for (size_t vi = 0; vi < videos.size(); vi++) {
        std::string verr;
        std::vector<uint8_t> md5s;
        const bool want_md5 = vi == 0 && !framemd5_path.empty();        // a framemd5 output without -map takes the first video stream
        if (want_md5) md5s.assign(16 * videos[vi].files.size(), 0);
        const int rc = encode_video_stream(videos[vi], encs[vi], devices, frames_in_flight, nthreads, slices, context, slicecrc, mux, flush_audio_until, pc, &verr,
                                           want_md5 ? &md5s : nullptr);
        if (!rc && want_md5) {
            // the file libavformat's framehash muxer writes (framehash.c ff_framehash_write_header, hashenc.c framehash_write_packet):
            // one rawvideo stream, time base = 1 / frame rate, pts = frame index
            const VideoStream& v = videos[vi];
            FILE* f = fopen(framemd5_path.c_str(), "w");
            if (!f) { cleanup(); return fail("cannot write " + framemd5_path, B200_ERR_IO); }
            fprintf(f, "#format: frame checksums\n#version: 2\n#hash: MD5\n#software: b200enc %u.%u.%u\n", b200_version() >> 16, (b200_version() >> 8) & 255, b200_version() & 255);
            fprintf(f, "#tb 0: %lld/%lld\n#media_type 0: video\n#codec_id 0: rawvideo\n#dimensions 0: %ux%u\n#sar 0: 0/1\n", v.fps_den, v.fps_num, v.info.width, v.info.height);
            fprintf(f, "#stream#, dts,        pts, duration,     size, hash\n");
            const size_t raw = b200_rawvideo_bytes(v.info.width, v.info.height, v.info.layout);
            for (size_t i = 0; i < v.files.size(); i++) {
                fprintf(f, "%d, %10lld, %10lld, %8lld, %8zu, ", 0, (long long)i, (long long)i, 1ll, raw);
                for (int k = 0; k < 16; k++) fprintf(f, "%02x", md5s[16 * i + k]);
                fputc('\n', f);
            }
            if (fclose(f)) { cleanup(); return fail("cannot write " + framemd5_path, B200_ERR_IO); }
        }
        encs[vi] = nullptr;                    // closed by the workers
        if (rc) { cleanup(); return fail(verr, rc); }
    }
    if (!flush_audio_until(INT64_MAX) || !mux.close()) { cleanup(); return fail(mux.error(), B200_ERR_IO); }
    pc.mark("mux close");
    cleanup();
    pc.mark("encoder close");
    return 0;
}
