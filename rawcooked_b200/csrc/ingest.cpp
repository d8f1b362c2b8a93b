#include "ingest.h"

#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../include/b200enc.h"
#include "ffv1_host.h"

namespace b200 {
namespace {
inline uint32_t rd32(const uint8_t* p, bool be) { return be ? ((uint32_t)p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3]) : ((uint32_t)p[3] << 24 | p[2] << 16 | p[1] << 8 | p[0]); }
inline uint16_t rd16(const uint8_t* p, bool be) { return be ? (uint16_t)(p[0] << 8 | p[1]) : (uint16_t)(p[1] << 8 | p[0]); }
bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
}  // namespace

char sniff(const uint8_t* h, size_t n) {
    if (n >= 4 && (!memcmp(h, "SDPX", 4) || !memcmp(h, "XPDS", 4))) return 'd';
    if (n >= 4 && ((h[0] == 'I' && h[1] == 'I' && h[2] == 42 && h[3] == 0) || (h[0] == 'M' && h[1] == 'M' && h[2] == 0 && h[3] == 42))) return 't';
    if (n >= 12 && (!memcmp(h, "RIFF", 4) || !memcmp(h, "RF64", 4)) && !memcmp(h + 8, "WAVE", 4)) return 'w';
    return 0;
}

bool parse_dpx(const uint8_t* h, size_t n, uint64_t file_size, ImageInfo* out, std::string* err) {
    if (n < 1664) { *err = "DPX header truncated"; return false; }
    bool be;
    if (!memcmp(h, "SDPX", 4)) be = true; else if (!memcmp(h, "XPDS", 4)) be = false; else { *err = "not a DPX file"; return false; }
    const uint32_t off_image = rd32(h + 4, be);
    if (rd16(h + 770, be) != 1) { *err = "DPX: more than one image element"; return false; }
    out->width = rd32(h + 772, be);
    out->height = rd32(h + 776, be);
    const uint8_t descriptor = h[800], bits = h[803];
    const uint16_t packing = rd16(h + 804, be), encoding = rd16(h + 806, be);
    uint32_t off_data = rd32(h + 808, be);
    if (!off_data) off_data = off_image;
    if (encoding) { *err = "DPX: RLE encoding is not supported"; return false; }
    if (descriptor != 50) { *err = "DPX: only RGB (descriptor 50) is supported by this encoder build"; return false; }
    int layout = -1;
    switch (bits) {
        case 8: layout = B200_DPX_RGB_8; break;
        case 10: if (packing == 1) layout = be ? B200_DPX_RGB_10_FILLED_A_BE : B200_DPX_RGB_10_FILLED_A_LE; break;
        case 12:
            if (packing == 1) layout = be ? B200_DPX_RGB_12_FILLED_A_BE : B200_DPX_RGB_12_FILLED_A_LE;
            else if (packing == 0 && be) layout = B200_DPX_RGB_12_PACKED_BE;
            break;
        case 16: layout = be ? B200_DPX_RGB_16_BE : B200_DPX_RGB_16_LE; break;
    }
    if (layout < 0) { *err = "DPX: unsupported bit depth / packing / endianness"; return false; }
    out->layout = layout;
    out->data_offset = off_data;
    out->data_bytes = layout_row_bytes(out->width, layout) * out->height;
    if (out->data_offset + out->data_bytes > file_size) { *err = "DPX: image data runs past the end of the file"; return false; }
    return true;
}

bool parse_tiff(const uint8_t* h, size_t n, uint64_t file_size, ImageInfo* out, std::string* err) {
    if (n < 8) { *err = "TIFF header truncated"; return false; }
    const bool be = h[0] == 'M';
    const uint32_t ifd = rd32(h + 4, be);
    if ((uint64_t)ifd + 2 > n) { *err = "TIFF: IFD outside the probed header"; return false; }
    const uint16_t cnt = rd16(h + ifd, be);
    if ((uint64_t)ifd + 2 + (uint64_t)cnt * 12 > n) { *err = "TIFF: IFD outside the probed header"; return false; }
    uint32_t w = 0, hgt = 0, bps = 0, compression = 1, photometric = 2, strip_off = 0, spp = 1, strip_bytes = 0, planar = 1, nstrips = 1;
    for (int i = 0; i < cnt; i++) {
        const uint8_t* e = h + ifd + 2 + i * 12;
        const uint16_t tag = rd16(e, be), type = rd16(e + 2, be);
        const uint32_t count = rd32(e + 4, be);
        auto val = [&]() -> uint32_t { return type == 3 ? rd16(e + 8, be) : rd32(e + 8, be); };
        switch (tag) {
            case 256: w = val(); break;
            case 257: hgt = val(); break;
            case 258:
                if (count == 1) bps = val();
                else { uint32_t o = rd32(e + 8, be); if ((uint64_t)o + 2 <= n) bps = rd16(h + o, be); }
                break;
            case 259: compression = val(); break;
            case 262: photometric = val(); break;
            case 273: nstrips = count; strip_off = count == 1 ? val() : 0; break;
            case 277: spp = val(); break;
            case 279: strip_bytes = count == 1 ? val() : 0; break;
            case 284: planar = val(); break;
        }
    }
    if (compression != 1 || photometric != 2 || spp != 3 || planar != 1 || nstrips != 1) { *err = "TIFF: only uncompressed single-strip chunky RGB is supported"; return false; }
    int layout = bps == 8 ? B200_TIFF_RGB_8 : bps == 16 ? (be ? B200_TIFF_RGB_16_BE : B200_TIFF_RGB_16_LE) : -1;
    if (layout < 0) { *err = "TIFF: unsupported bit depth"; return false; }
    out->width = w; out->height = hgt; out->layout = layout; out->data_offset = strip_off;
    out->data_bytes = layout_row_bytes(w, layout) * hgt;
    if (strip_bytes && strip_bytes != out->data_bytes) { *err = "TIFF: strip byte count does not match the image size"; return false; }
    if (out->data_offset + out->data_bytes > file_size) { *err = "TIFF: image data runs past the end of the file"; return false; }
    return true;
}

bool parse_wav(const uint8_t* h, size_t n, uint64_t file_size, WavInfo* out, std::string* err) {
    if (n < 12 || sniff(h, n) != 'w') { *err = "not a RIFF/WAVE file"; return false; }
    const bool rf64 = !memcmp(h, "RF64", 4);
    uint64_t pos = 12, ds64_data = 0;
    bool have_fmt = false;
    while (pos + 8 <= n) {
        const uint8_t* c = h + pos;
        uint64_t sz = rd32(c + 4, false);
        if (!memcmp(c, "ds64", 4) && pos + 8 + 16 <= n) {
            ds64_data = (uint64_t)rd32(c + 16, false) | ((uint64_t)rd32(c + 20, false) << 32);
        } else if (!memcmp(c, "fmt ", 4) && pos + 8 + 16 <= n) {
            const uint16_t tag = rd16(c + 8, false);
            out->channels = rd16(c + 10, false);
            out->sample_rate = rd32(c + 12, false);
            out->bits = rd16(c + 22, false);
            uint16_t fmt = tag;
            if (tag == 0xFFFE && sz >= 26 && pos + 8 + 26 <= n) fmt = rd16(c + 8 + 24, false);
            out->is_float = fmt == 3;
            if (fmt != 1 && fmt != 3) { *err = "WAV: unsupported format tag"; return false; }
            have_fmt = true;
        } else if (!memcmp(c, "data", 4)) {
            if (!have_fmt) { *err = "WAV: data chunk before fmt chunk"; return false; }
            if (rf64 && sz == 0xFFFFFFFFu) sz = ds64_data;
            out->data_offset = pos + 8;
            out->data_bytes = sz;
            if (out->data_offset + out->data_bytes > file_size) out->data_bytes = file_size - out->data_offset;   // truncated file
            return true;
        }
        pos += 8 + sz + (sz & 1);
    }
    *err = "WAV: data chunk not found in the probed header";
    return false;
}

// ffmpeg's image2 patterns (libavformat/utils av_get_frame_filename): "%d", "%0Nd" and "%%" only, exactly one number. The
// pattern comes from argv, so it is parsed by hand: it never reaches a printf-style function as the format.
static bool image2_name(const std::string& pattern, long long n, std::string* out) {
    out->clear();
    bool have_number = false;
    for (size_t i = 0; i < pattern.size(); i++) {
        const char c = pattern[i];
        if (c != '%') { out->push_back(c); continue; }
        if (i + 1 < pattern.size() && pattern[i + 1] == '%') { out->push_back('%'); i++; continue; }
        size_t j = i + 1;
        int width = 0;
        const bool zero = j < pattern.size() && pattern[j] == '0';
        while (j < pattern.size() && pattern[j] >= '0' && pattern[j] <= '9') { width = width * 10 + (pattern[j] - '0'); if (width > 64) return false; j++; }
        if (j >= pattern.size() || pattern[j] != 'd' || have_number) return false;     // any other conversion is refused
        std::string digits = std::to_string(n < 0 ? -n : n);
        if ((int)digits.size() < width) digits.insert(0, (size_t)width - digits.size(), zero ? '0' : ' ');
        if (n < 0) digits.insert(0, "-");
        *out += digits;
        have_number = true;
        i = j;
    }
    return have_number;
}

std::vector<std::string> expand_image2(const std::string& pattern, long long start) {
    std::vector<std::string> out;
    if (pattern.find('%') == std::string::npos) { if (file_exists(pattern)) out.push_back(pattern); return out; }
    std::string name;
    for (long long n = start;; n++) {
        if (!image2_name(pattern, n, &name) || !file_exists(name)) break;
        out.push_back(name);
    }
    return out;
}

std::vector<std::string> read_concat_list(const std::string& list_path) {
    std::vector<std::string> out;
    std::ifstream f(list_path);
    std::string dir;
    const size_t sl = list_path.find_last_of('/');
    if (sl != std::string::npos) dir = list_path.substr(0, sl + 1);
    std::string line;
    while (std::getline(f, line)) {
        if (line.compare(0, 5, "file ") != 0) continue;
        std::string p = line.substr(5);
        while (!p.empty() && (p.back() == '\r' || p.back() == ' ')) p.pop_back();
        if (p.size() >= 2 && p.front() == '\'' && p.back() == '\'') p = p.substr(1, p.size() - 2);
        size_t k;
        while ((k = p.find("'\\''")) != std::string::npos) p.replace(k, 4, "'");   // ffmpeg's quoting of a single quote
        if (!p.empty() && p[0] != '/') p = dir + p;
        out.push_back(p);
    }
    return out;
}

}  // namespace b200
