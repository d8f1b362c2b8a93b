#include "mkv_mux.h"

#include <fcntl.h>
#include <unistd.h>

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

}  // namespace

MkvWriter::~MkvWriter() { if (fd_ >= 0) ::close(fd_); }

// Packets are tens of megabytes: they go straight from the caller's buffer to write(2); only the small pieces (element
// headers) are gathered in a buffer. Sizes that are not known when an element starts (Segment, Cluster) are written as
// 8-byte placeholders and patched with pwrite once the element is complete.
bool MkvWriter::flush_small() {
    size_t done = 0;
    while (done < small_.size()) {
        const ssize_t r = ::write(fd_, small_.data() + done, small_.size() - done);
        if (r <= 0) { err_ = "write failed"; return false; }
        done += (size_t)r;
    }
    small_.clear();
    return true;
}

bool MkvWriter::put(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    if (n < (256u << 10)) {
        small_.insert(small_.end(), b, b + n);
        if (small_.size() > (4u << 20) && !flush_small()) return false;
    } else {
        if (!flush_small()) return false;
        size_t done = 0;
        while (done < n) {
            const ssize_t r = ::write(fd_, b + done, n - done);
            if (r <= 0) { err_ = "write failed"; return false; }
            done += (size_t)r;
        }
    }
    pos_ += n;
    return true;
}

bool MkvWriter::patch_size8(uint64_t at, uint64_t value) {
    if (!flush_small()) return false;
    std::vector<uint8_t> sz;
    put_size8(sz, value);
    if (::pwrite(fd_, sz.data(), 8, (off_t)at) != 8) { err_ = "cannot patch an element size"; return false; }
    return true;
}

bool MkvWriter::open(const std::string& path, const std::vector<MkvTrack>& tracks, const std::vector<MkvAttachment>& atts, double duration_ms) {
    fd_ = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd_ < 0) { err_ = "cannot create " + path; return false; }
    std::vector<uint8_t> h, body;
    el_uint(body, 0x4286, 1); el_uint(body, 0x42F7, 1); el_uint(body, 0x42F2, 4); el_uint(body, 0x42F3, 8);
    el_str(body, 0x4282, "matroska"); el_uint(body, 0x4287, 4); el_uint(body, 0x4285, 2);
    el_master(h, 0x1A45DFA3, body);
    put_id(h, 0x18538067);                                   // Segment, size patched in close()
    segment_size_pos_ = h.size();
    put_size8(h, 0);
    if (!put(h.data(), h.size())) return false;
    segment_data_start_ = pos_;

    std::vector<uint8_t> seg;
    {   // Info
        std::vector<uint8_t> info;
        el_uint(info, 0x2AD7B1, 1000000);                    // TimestampScale: 1 ms
        el_str(info, 0x4D80, "b200enc (RAWcooked B200 encode path)");
        el_str(info, 0x5741, "b200enc");
        if (duration_ms > 0) el_float(info, 0x4489, duration_ms);
        el_master(seg, 0x1549A966, info);
    }
    {   // Tracks: TrackNumber = 1-based order, which the reference requires of the SimpleBlock track vint (Matroska.cpp:938-942)
        std::vector<uint8_t> trs;
        int num = 0;
        for (const MkvTrack& t : tracks) {
            num++;
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
        put_id(h, 0x1F43B675);
        cluster_size_at_ = pos_ + h.size();
        put_size8(h, 0);
        cluster_data_start_ = pos_ + h.size();
        el_uint(h, 0xE7, (uint64_t)cluster_time_);
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
    ok = ok && patch_size8(segment_data_start_ - 8, pos_ - segment_data_start_);
    if (::close(fd_) != 0 && ok) { err_ = "close failed"; ok = false; }
    fd_ = -1;
    return ok;
}

}  // namespace b200
