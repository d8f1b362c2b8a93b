#include "mkv_mux.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>

namespace b200 {
namespace {

void put_id(std::vector<uint8_t>& o, uint32_t id) {
    if (id > 0xFFFFFF) o.push_back((uint8_t)(id >> 24));
    if (id > 0xFFFF) o.push_back((uint8_t)(id >> 16));
    if (id > 0xFF) o.push_back((uint8_t)(id >> 8));
    o.push_back((uint8_t)id);
}
void put_size(std::vector<uint8_t>& o, uint64_t n) {         // shortest EBML vint that is not the reserved all-ones value
    int len = 1;
    while (len < 8 && n >= ((uint64_t)1 << (7 * len)) - 1) len++;
    for (int i = len - 1; i >= 0; i--) {
        uint8_t b = (uint8_t)(n >> (8 * i));
        if (i == len - 1) b |= (uint8_t)(0x80 >> (len - 1));
        o.push_back(b);
    }
}
void put_size8(std::vector<uint8_t>& o, uint64_t n) {        // fixed 8-byte size (patched later)
    o.push_back(0x01);
    for (int i = 6; i >= 0; i--) o.push_back((uint8_t)(n >> (8 * i)));
}
void el_bytes(std::vector<uint8_t>& o, uint32_t id, const void* p, size_t n) {
    put_id(o, id); put_size(o, n);
    const uint8_t* b = static_cast<const uint8_t*>(p);
    o.insert(o.end(), b, b + n);
}
void el_uint(std::vector<uint8_t>& o, uint32_t id, uint64_t v) {
    uint8_t tmp[8]; int n = 1;
    while (n < 8 && (v >> (8 * n))) n++;
    for (int i = 0; i < n; i++) tmp[i] = (uint8_t)(v >> (8 * (n - 1 - i)));
    el_bytes(o, id, tmp, n);
}
void el_uint_fixed(std::vector<uint8_t>& o, uint32_t id, uint64_t v, int n) {
    uint8_t tmp[8];
    for (int i = 0; i < n; i++) tmp[i] = (uint8_t)(v >> (8 * (n - 1 - i)));
    el_bytes(o, id, tmp, n);
}
void el_str(std::vector<uint8_t>& o, uint32_t id, const std::string& s) { el_bytes(o, id, s.data(), s.size()); }
void el_float(std::vector<uint8_t>& o, uint32_t id, double d) {
    uint64_t u; std::memcpy(&u, &d, 8);
    uint8_t tmp[8];
    for (int i = 0; i < 8; i++) tmp[i] = (uint8_t)(u >> (8 * (7 - i)));
    el_bytes(o, id, tmp, 8);
}
void el_master(std::vector<uint8_t>& o, uint32_t id, const std::vector<uint8_t>& body) { el_bytes(o, id, body.data(), body.size()); }
constexpr size_t kSeekHeadRoom = 128;

}  // namespace

MkvWriter::~MkvWriter() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_job_.notify_all();
    for (auto& t : writers_) if (t.joinable()) t.join();
    if (fd_ >= 0) ::close(fd_);
}

bool MkvWriter::pwrite_all(const void* p, size_t n, uint64_t at) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    size_t done = 0;
    while (done < n) {
        const ssize_t r = ::pwrite(fd_, b + done, n - done, (off_t)(at + done));
        if (r <= 0) return false;
        done += (size_t)r;
    }
    return true;
}

// One large piece of a packet. On tmpfs, writes to one file serialise on the inode lock however many threads issue them; there
// the piece goes through a private mapping of its page range instead (the page faults of different threads run in parallel).
bool MkvWriter::write_piece(const uint8_t* p, size_t n, uint64_t at) {
    if (!use_mmap_) return pwrite_all(p, n, at);
    const uint64_t page = 4096, a0 = at & ~(page - 1), a1 = (at + n + page - 1) & ~(page - 1);
    void* m = ::mmap(nullptr, (size_t)(a1 - a0), PROT_READ | PROT_WRITE, MAP_SHARED, fd_, (off_t)a0);
    if (m == MAP_FAILED) return pwrite_all(p, n, at);
    std::memcpy(static_cast<uint8_t*>(m) + (at - a0), p, n);
    ::munmap(m, (size_t)(a1 - a0));
    return true;
}

void MkvWriter::writer_loop() {
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_job_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
            if (jobs_.empty()) return;
            j = jobs_.front();
            jobs_.pop_front();
        }
        const bool ok = write_piece(j.p, j.n, j.at);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) io_failed_ = true;
            inflight_--;
        }
        cv_done_.notify_all();
    }
}

bool MkvWriter::sync() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return inflight_ == 0; });
    if (io_failed_) { err_ = "write failed"; return false; }
    return true;
}

// Packets are tens of megabytes: they go straight from the caller's buffer to pwrite(2) on the writer threads; only the small
// pieces (element headers) are gathered in a buffer. Sizes that are not known when an element starts (Segment, Cluster) are
// written as 8-byte placeholders and patched once the element is complete.
bool MkvWriter::flush_small() {
    if (small_.empty()) return true;
    if (!pwrite_all(small_.data(), small_.size(), small_at_)) { err_ = "write failed"; return false; }
    small_.clear();
    return true;
}

bool MkvWriter::put(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    if (n < (256u << 10)) {
        if (small_.empty()) small_at_ = pos_;
        small_.insert(small_.end(), b, b + n);
        pos_ += n;
        if (small_.size() > (4u << 20) && !flush_small()) return false;
        return true;
    }
    if (!flush_small()) return false;
    if (use_mmap_ && pos_ + n > file_size_) {      // a mapping cannot extend the file: grow it ahead of the writers, trim in close()
        file_size_ = ((pos_ + n) | (((uint64_t)1 << 30) - 1)) + 1;
        if (::ftruncate(fd_, (off_t)file_size_) != 0) { use_mmap_ = false; }
    }
    // large payload: cut into pieces for the writer pool
    const size_t piece = (size_t)8 << 20;
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (io_failed_) { err_ = "write failed"; return false; }
        for (size_t o = 0; o < n; o += piece) {
            jobs_.push_back({b + o, n - o < piece ? n - o : piece, pos_ + o});
            inflight_++;
        }
    }
    cv_job_.notify_all();
    pos_ += n;
    return true;
}

bool MkvWriter::patch_size8(uint64_t at, uint64_t value) {
    std::vector<uint8_t> sz;
    put_size8(sz, value);
    if (!small_.empty() && at >= small_at_ && at + 8 <= small_at_ + small_.size()) {     // still in the header buffer
        std::memcpy(small_.data() + (at - small_at_), sz.data(), 8);
        return true;
    }
    if (!pwrite_all(sz.data(), 8, at)) { err_ = "cannot patch an element size"; return false; }
    return true;
}

bool MkvWriter::open(const std::string& path, const std::vector<MkvTrack>& tracks, const std::vector<MkvAttachment>& atts, double duration_ms) {
    fd_ = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd_ < 0) { err_ = "cannot create " + path; return false; }
    {
        struct statfs fs;
        use_mmap_ = ::fstatfs(fd_, &fs) == 0 && (unsigned long)fs.f_type == 0x01021994ul && !getenv("B200_NO_MMAP_OUTPUT");   // TMPFS_MAGIC
        unsigned nw = std::thread::hardware_concurrency();
        nw = nw < 2 ? 2 : nw > 16 ? 16 : nw;
        for (unsigned i = 0; i < nw; i++) writers_.emplace_back([this] { writer_loop(); });
    }
    std::vector<uint8_t> h, body;
    el_uint(body, 0x4286, 1); el_uint(body, 0x42F7, 1); el_uint(body, 0x42F2, 4); el_uint(body, 0x42F3, 8);
    el_str(body, 0x4282, "matroska"); el_uint(body, 0x4287, 4); el_uint(body, 0x4285, 2);
    el_master(h, 0x1A45DFA3, body);
    put_id(h, 0x18538067);                                   // Segment, size patched in close()
    put_size8(h, 0);
    if (!put(h.data(), h.size())) return false;
    segment_data_start_ = pos_;

    std::vector<uint8_t> seg;
    {   // room for the SeekHead (Info, Tracks, Attachments, Cues: 4 entries of at most 21 bytes + header), a Void until close()
        seekhead_at_ = pos_;
        put_id(seg, 0xEC);
        put_size(seg, kSeekHeadRoom - 2);
        seg.resize(seg.size() + kSeekHeadRoom - 2, 0);
    }
    info_at_ = pos_ + seg.size();
    {   // Info
        std::vector<uint8_t> info;
        el_uint(info, 0x2AD7B1, 1000000);                    // TimestampScale: 1 ms
        el_str(info, 0x4D80, "b200enc (RAWcooked B200 encode path)");
        el_str(info, 0x5741, "b200enc");
        if (duration_ms > 0) el_float(info, 0x4489, duration_ms);
        el_master(seg, 0x1549A966, info);
    }
    tracks_at_ = pos_ + seg.size();
    {   // Tracks: TrackNumber = 1-based order, which the reference requires of the SimpleBlock track vint (Matroska.cpp:938-942)
        std::vector<uint8_t> trs;
        int num = 0;
        for (const MkvTrack& t : tracks) {
            num++;
            if (t.video && !video_track_) video_track_ = num;
            std::vector<uint8_t> te;
            el_uint(te, 0xD7, (uint64_t)num);
            el_uint(te, 0x73C5, (uint64_t)num);
            el_uint(te, 0x83, t.video ? 1 : 2);
            el_uint(te, 0x9C, 0);                            // FlagLacing
            el_str(te, 0x86, t.codec_id);
            if (!t.codec_private.empty()) el_bytes(te, 0x63A2, t.codec_private.data(), t.codec_private.size());
            if (t.video) {
                if (t.frame_rate > 0) el_uint(te, 0x23E383, (uint64_t)(1e9 / t.frame_rate + 0.5));   // DefaultDuration
                std::vector<uint8_t> v;
                el_uint(v, 0x9A, 2);                         // FlagInterlaced: progressive
                el_uint_fixed(v, 0xB0, t.width, 2);          // the reference reads these only when 1-2 bytes long (:1007-1030)
                el_uint_fixed(v, 0xBA, t.height, 2);
                el_master(te, 0xE0, v);
            } else {
                std::vector<uint8_t> a;
                el_float(a, 0xB5, (double)t.sample_rate);
                el_uint(a, 0x9F, t.channels);
                el_uint(a, 0x6264, t.bit_depth);
                el_master(te, 0xE1, a);
            }
            el_master(trs, 0xAE, te);
        }
        el_master(seg, 0x1654AE6B, trs);
    }
    attachments_at_ = atts.empty() ? 0 : pos_ + seg.size();
    if (!atts.empty()) {   // Attachments must precede the first Cluster for the reference to find the sidecar (Matroska.cpp:861-874)
        std::vector<uint8_t> all;
        uint64_t uid = 0x5241574330000001ull;
        for (const MkvAttachment& a : atts) {
            std::vector<uint8_t> af;
            el_str(af, 0x466E, a.name);
            el_str(af, 0x4660, a.mime.empty() ? "application/octet-stream" : a.mime);
            el_bytes(af, 0x465C, a.data.data(), a.data.size());
            el_uint_fixed(af, 0x46AE, uid++, 8);
            el_master(all, 0x61A7, af);
        }
        el_master(seg, 0x1941A469, all);
    }
    return put(seg.data(), seg.size());
}

bool MkvWriter::flush_cluster() {
    if (cluster_time_ < 0) return true;
    const bool ok = patch_size8(cluster_size_at_, pos_ - cluster_data_start_);
    cluster_time_ = -1;
    return ok;
}

bool MkvWriter::write_block(int track, int64_t time_ms, const uint8_t* data, size_t len, bool keyframe) {
    if (track < 1 || track > 126) { err_ = "bad track number"; return false; }
    // new Cluster at most every 5 s / 32 MiB, and always before the 16-bit relative timestamp would overflow
    if (cluster_time_ >= 0 && (time_ms - cluster_time_ > 5000 || time_ms < cluster_time_ || pos_ - cluster_data_start_ > (32u << 20)))
        if (!flush_cluster()) return false;
    std::vector<uint8_t> h;
    if (cluster_time_ < 0) {
        cluster_time_ = time_ms;
        cluster_has_cue_ = false;
        cluster_pos_ = pos_;
        put_id(h, 0x1F43B675);
        cluster_size_at_ = pos_ + h.size();
        put_size8(h, 0);
        cluster_data_start_ = pos_ + h.size();
        el_uint(h, 0xE7, (uint64_t)cluster_time_);
    }
    if (!cluster_has_cue_ && (video_track_ == 0 || track == video_track_)) {     // one CuePoint per Cluster, on the first video track
        cues_.push_back({time_ms, track, cluster_pos_ - segment_data_start_});
        cluster_has_cue_ = true;
    }
    const int64_t rel = time_ms - cluster_time_;
    put_id(h, 0xA3);
    put_size(h, len + 4);
    h.push_back((uint8_t)(0x80 | track));
    h.push_back((uint8_t)(rel >> 8));
    h.push_back((uint8_t)rel);
    h.push_back(keyframe ? 0x80 : 0x00);
    return put(h.data(), h.size()) && put(data, len);
}

bool MkvWriter::close() {
    if (fd_ < 0) return true;
    bool ok = flush_cluster();
    uint64_t cues_at = 0;
    if (ok && !cues_.empty()) {   // Cues: CuePoint{CueTime, CueTrackPositions{CueTrack, CueClusterPosition}} per Cluster
        std::vector<uint8_t> all;
        for (const Cue& c : cues_) {
            std::vector<uint8_t> cp, tp;
            el_uint(cp, 0xB3, (uint64_t)c.time);
            el_uint(tp, 0xF7, (uint64_t)c.track);
            el_uint(tp, 0xF1, c.cluster_pos);
            el_master(cp, 0xB7, tp);
            el_master(all, 0xBB, cp);
        }
        std::vector<uint8_t> cues;
        el_master(cues, 0x1C53BB6B, all);
        cues_at = pos_;
        ok = put(cues.data(), cues.size());
    }
    if (ok) {                     // SeekHead into the room reserved at the start of the Segment, the rest stays a Void
        std::vector<uint8_t> sh_body;
        auto seek = [&](uint32_t id, uint64_t at) {
            if (!at) return;
            std::vector<uint8_t> e, idb;
            put_id(idb, id);
            el_bytes(e, 0x53AB, idb.data(), idb.size());
            el_uint_fixed(e, 0x53AC, at - segment_data_start_, 8);
            el_master(sh_body, 0x4DBB, e);
        };
        seek(0x1549A966, info_at_); seek(0x1654AE6B, tracks_at_); seek(0x1941A469, attachments_at_); seek(0x1C53BB6B, cues_at);
        std::vector<uint8_t> sh;
        el_master(sh, 0x114D9B74, sh_body);
        if (sh.size() + 2 <= kSeekHeadRoom) {
            const size_t rest = kSeekHeadRoom - sh.size();
            put_id(sh, 0xEC);
            put_size(sh, rest - 2);
            sh.resize(kSeekHeadRoom, 0);
            ok = flush_small() && pwrite_all(sh.data(), sh.size(), seekhead_at_);
            if (!ok && err_.empty()) err_ = "write failed";
        }
    }
    ok = ok && patch_size8(segment_data_start_ - 8, pos_ - segment_data_start_);
    ok = flush_small() && ok;
    ok = sync() && ok;
    if (use_mmap_ && file_size_ > pos_ && ::ftruncate(fd_, (off_t)pos_) != 0 && ok) { err_ = "cannot trim the output file"; ok = false; }
    if (::close(fd_) != 0 && ok) { err_ = "close failed"; ok = false; }
    fd_ = -1;
    return ok;
}

}  // namespace b200
