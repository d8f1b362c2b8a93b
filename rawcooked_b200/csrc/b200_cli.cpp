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
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200enc.h"
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

struct VideoStream {
    std::vector<std::string> files;
    b200::ImageInfo info;
    double fps = 24;
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

int b200_flac_encode_file_to_mux(const std::string& path, const b200::WavInfo& wi, int track, int device,
                                 std::vector<uint8_t>* codec_private, std::vector<std::pair<int64_t, std::vector<uint8_t>>>* packets,
                                 std::string* err);   // flac_host.cpp

extern "C" int b200enc_main(int argc, char** argv) {
    PhaseClock pc;
    std::vector<InputSpec> inputs;
    std::vector<AttachSpec> attaches;
    std::map<std::string, std::string> pending, outopt;
    std::vector<std::string> outputs;
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
        // a bare token is an output file; the options gathered since the last -i / output belong to it
        for (auto& kv : pending) outopt[kv.first] = kv.second;
        pending.clear();
        outputs.push_back(a);
    }
    if (inputs.empty()) return fail("no input (-i)");
    if (outputs.empty()) return fail("no output file");
    if (outputs.size() > 1) return fail("only one Matroska output is supported (framemd5 and extra outputs are not)");
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
    int device = 0;
    if (const char* e = getenv("B200_DEVICE")) device = atoi(e);

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

    // ---- encode: video in batches through the GPU, audio packets interleaved by timestamp
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
        VideoStream& v = videos[vi];
        b200_ffv1_enc* E = encs[vi];
        const size_t fb = b200_ffv1_frame_bytes(v.info.width, v.info.height, v.info.layout);
        const size_t B = std::min<size_t>((size_t)frames_in_flight, v.files.size());
        // Source-file ingest at rate: the payloads of a batch are read by `nread` threads (pread straight into pinned
        // memory, one header parse per file because the payload offset may differ from frame to frame), and the NEXT batch is
        // read while the GPU encodes the current one (two pinned input buffers).
        uint8_t* h_in[2] = {nullptr, nullptr};
        // packets leave the device one by one through a two-slot pinned ring (a packet is tens of MB; pinning a buffer for
        // a whole batch of worst-case packets would take longer than encoding it)
        uint8_t* h_ring[2] = {nullptr, nullptr};
        size_t ring_cap = std::max<size_t>(fb + (fb >> 2), (size_t)1 << 20);
        cudaStream_t cs = nullptr;
        cudaEvent_t cev[2] = {nullptr, nullptr};
        const bool two = v.files.size() > B;
        // pinning memory is slow (a few GB/s): the second input buffer is pinned in the background while the first batch is
        // read and encoded
        bool in1_ok = true;
        std::thread pin1;
        if (two) pin1 = std::thread([&] { cudaSetDevice(device); in1_ok = cudaHostAlloc((void**)&h_in[1], fb * B, cudaHostAllocDefault) == cudaSuccess; });
        if (cudaHostAlloc((void**)&h_in[0], fb * B, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&h_ring[0], ring_cap, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&h_ring[1], ring_cap, cudaHostAllocDefault) != cudaSuccess ||
            cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&cev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&cev[1], cudaEventDisableTiming) != cudaSuccess) {
            if (pin1.joinable()) pin1.join();
            cleanup(); return fail("cannot allocate pinned host buffers");
        }
        unsigned nread = std::thread::hardware_concurrency();
        if (outopt.count("-threads") && atoi(outopt["-threads"].c_str()) > 0) nread = (unsigned)atoi(outopt["-threads"].c_str());
        nread = std::max(1u, std::min(nread, 32u));
        // reads frames [f0, f0 + n) into dst; returns "" or the first error
        auto read_batch = [&](size_t f0, size_t n, uint8_t* dst) -> std::string {
            std::vector<std::string> errs(nread);
            auto work = [&](unsigned t) {
                for (size_t k = t; k < n && errs[t].empty(); k += nread) {
                    const std::string& path = v.files[f0 + k];
                    const int fd = open(path.c_str(), O_RDONLY);
                    if (fd < 0) { errs[t] = "cannot open " + path; return; }
                    b200::ImageInfo fi = v.info;
                    if (f0 + k > 0) {
                        std::vector<uint8_t> head(65536);
                        const ssize_t hn = pread(fd, head.data(), head.size(), 0);
                        std::string err;
                        const bool ok = v.kind == 'd' ? b200::parse_dpx(head.data(), hn > 0 ? (size_t)hn : 0, file_size(path), &fi, &err)
                                                      : b200::parse_tiff(head.data(), hn > 0 ? (size_t)hn : 0, file_size(path), &fi, &err);
                        if (!ok || fi.width != v.info.width || fi.height != v.info.height || fi.layout != v.info.layout) {
                            close(fd);
                            errs[t] = path + ": " + (ok ? std::string("geometry differs from the first frame") : err);
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
        std::vector<const uint8_t*> ptrs(B);
        std::vector<size_t> off(B), len(B);
        pc.mark("pinned buffers");
        std::string rerr = read_batch(0, std::min(B, v.files.size()), h_in[0]);
        pc.mark("first batch read");
        if (!rerr.empty()) { cleanup(); return fail(rerr, B200_ERR_IO); }
        int cur = 0;
        for (size_t f0 = 0; f0 < v.files.size(); f0 += B) {
            const size_t n = std::min(B, v.files.size() - f0);
            std::string next_err;
            std::thread prefetch;
            if (pin1.joinable()) { pin1.join(); if (!in1_ok) { cleanup(); return fail("cannot allocate pinned host buffers"); } }
            if (f0 + B < v.files.size())
                prefetch = std::thread([&, f0] { next_err = read_batch(f0 + B, std::min(B, v.files.size() - f0 - B), h_in[cur ^ 1]); });
            for (size_t k = 0; k < n; k++) ptrs[k] = h_in[cur] + k * fb;
            int rc = b200_ffv1_submit_host(E, ptrs.data(), (int32_t)n);
            const void* d_arena = nullptr;
            if (!rc) rc = b200_ffv1_packets_device(E, &d_arena, off.data(), len.data(), (int32_t)n);   // waits for the encode
            pc.mark("encode (batch)");
            bool mux_ok = rc == 0;
            // packet k+1 crosses PCIe while packet k is written to the file
            auto fetch = [&](size_t k) -> bool {
                const int sl = (int)(k & 1);
                if (len[k] > ring_cap) {                       // a packet larger than 1.25 x the frame: grow the ring
                    cudaStreamSynchronize(cs);
                    cudaFreeHost(h_ring[0]); cudaFreeHost(h_ring[1]);
                    ring_cap = len[k] + (len[k] >> 3);
                    if (cudaHostAlloc((void**)&h_ring[0], ring_cap, cudaHostAllocDefault) != cudaSuccess ||
                        cudaHostAlloc((void**)&h_ring[1], ring_cap, cudaHostAllocDefault) != cudaSuccess) return false;
                }
                return cudaMemcpyAsync(h_ring[sl], static_cast<const uint8_t*>(d_arena) + off[k], len[k], cudaMemcpyDeviceToHost, cs) == cudaSuccess &&
                       cudaEventRecord(cev[sl], cs) == cudaSuccess;
            };
            if (mux_ok) mux_ok = fetch(0);
            for (size_t k = 0; k < n && mux_ok; k++) {
                if (cudaEventSynchronize(cev[k & 1]) != cudaSuccess) { mux_ok = false; break; }
                // the ring may only be re-used (or re-allocated) once the block that lives in it has been written
                const int64_t t_ms = (int64_t)std::llround(1000.0 * (double)(f0 + k) / v.fps);
                if (k + 1 < n && len[k + 1] <= ring_cap && !fetch(k + 1)) { mux_ok = false; break; }
                mux_ok = flush_audio_until(t_ms) && mux.write_block(v.track, t_ms, h_ring[k & 1], len[k]);
                if (mux_ok && k + 1 < n && len[k + 1] > ring_cap) mux_ok = fetch(k + 1);
            }
            pc.mark("mux write (batch)");
            if (prefetch.joinable()) prefetch.join();
            pc.mark("wait for next batch read");
            if (rc) { cleanup(); return fail(std::string("ffv1 encode: ") + b200_last_error()); }
            if (!mux_ok) { cleanup(); return fail(mux.error(), B200_ERR_IO); }
            if (!next_err.empty()) { cleanup(); return fail(next_err, B200_ERR_IO); }
            cur ^= 1;
        }
        cudaStreamSynchronize(cs);
        cudaFreeHost(h_in[0]);
        if (h_in[1]) cudaFreeHost(h_in[1]);
        cudaFreeHost(h_ring[0]); cudaFreeHost(h_ring[1]);
        cudaEventDestroy(cev[0]); cudaEventDestroy(cev[1]);
        cudaStreamDestroy(cs);
    }
    if (!flush_audio_until(INT64_MAX) || !mux.close()) { cleanup(); return fail(mux.error(), B200_ERR_IO); }
    pc.mark("mux close");
    cleanup();
    pc.mark("encoder close");
    return 0;
}
