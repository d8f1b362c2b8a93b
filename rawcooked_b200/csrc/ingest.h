// Payload locators for the source files the encoder front-end re-opens itself (the reference only hands over paths:
// /root/reference/Source/CLI/Output.cpp:116-136, 227-250). Restates just enough of the reference's parsers to find the
// image/audio payload and its layout: DPX header (Source/Lib/Uncompressed/DPX/DPX.cpp:250-416, flavors :184-231),
// baseline TIFF (TIFF/TIFF.cpp:380-717), RIFF/WAVE (WAV/WAV.cpp:271-436).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct ImageInfo {
    uint32_t width = 0, height = 0;
    int layout = -1;            // b200_layout
    uint64_t data_offset = 0;   // first payload byte in the file
    uint64_t data_bytes = 0;    // payload length
};

struct WavInfo {
    uint32_t sample_rate = 0, channels = 0, bits = 0;
    uint64_t data_offset = 0, data_bytes = 0;
    bool is_float = false;
};

// `head` = first bytes of the file (>= 2048 for DPX; whole IFD area for TIFF: pass up to 64 KiB), `file_size` = size on disk
bool parse_dpx(const uint8_t* head, size_t n, uint64_t file_size, ImageInfo* out, std::string* err);
bool parse_tiff(const uint8_t* head, size_t n, uint64_t file_size, ImageInfo* out, std::string* err);
bool parse_wav(const uint8_t* head, size_t n, uint64_t file_size, WavInfo* out, std::string* err);
// sniff by magic: 'd' DPX, 't' TIFF, 'w' WAV, 0 unknown
char sniff(const uint8_t* head, size_t n);

// ffmpeg image2 pattern ("%06d" style, exactly one conversion) + start number -> existing files in order
std::vector<std::string> expand_image2(const std::string& pattern, long long start_number);
// ffmpeg concat demuxer list (`file 'path'` lines) -> paths (relative to the list's directory)
std::vector<std::string> read_concat_list(const std::string& list_path);

}  // namespace b200
