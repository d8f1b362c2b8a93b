// Minimal Matroska writer for the encoder front-end: exactly what the reference's parser consumes in `--check`
// (/root/reference/Source/Lib/Compressed/Matroska/Matroska.cpp: element tables :128-217, SimpleBlock :934-953, TrackEntry
// :976-1030, attachments :523-595, finite Segment size :1259-1277) and what players need (Info, DefaultDuration, Audio).
// Layout: EBML | Segment(size patched at close){ Info, Tracks, Attachments (before the first Cluster), Cluster* }.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace b200 {

struct MkvTrack {
    bool video = true;
    std::string codec_id;                 // "V_FFV1", "A_FLAC", "A_PCM/INT/LIT"
    std::vector<uint8_t> codec_private;
    uint32_t width = 0, height = 0;       // video
    double frame_rate = 24.0;             // video
    uint32_t sample_rate = 0, channels = 0, bit_depth = 0;   // audio
};

struct MkvAttachment {
    std::string name, mime;
    std::vector<uint8_t> data;
};

class MkvWriter {
  public:
    ~MkvWriter();
    // returns false (and sets error()) on I/O failure
    bool open(const std::string& path, const std::vector<MkvTrack>& tracks, const std::vector<MkvAttachment>& attachments, double duration_ms);
    // track is 1-based TrackEntry order; time in milliseconds
    bool write_block(int track, int64_t time_ms, const uint8_t* data, size_t len, bool keyframe = true);
    bool close();
    const std::string& error() const { return err_; }
    uint64_t bytes_written() const { return pos_; }

  private:
    bool put(const void* p, size_t n);
    bool flush_small();
    bool patch_size8(uint64_t at, uint64_t value);
    bool flush_cluster();
    int fd_ = -1;
    std::string err_;
    uint64_t pos_ = 0, segment_size_pos_ = 0, segment_data_start_ = 0;
    std::vector<uint8_t> small_;          // element headers not yet written
    uint64_t cluster_size_at_ = 0, cluster_data_start_ = 0;   // file offsets of the open Cluster's size field / first child
    int64_t cluster_time_ = -1;
};

}  // namespace b200
