// Minimal Matroska writer for the encoder front-end: exactly what the reference's parser consumes in `--check`
// (/root/reference/Source/Lib/Compressed/Matroska/Matroska.cpp: element tables :128-217, SimpleBlock :934-953, TrackEntry
// :976-1030, attachments :523-595, finite Segment size :1259-1277) and what players need (Info, DefaultDuration, Audio).
// Layout: EBML | Segment(size patched at close){ SeekHead (room reserved, filled at close), Info, Tracks, Attachments (before the
// first Cluster), Cluster*, Cues }.
// Every byte goes out with pwrite at an offset fixed when the block is queued, so the large packets are written by a small
// pool of threads in parallel while the caller prepares the next ones; sync() waits for them.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
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
    // waits until every block queued so far is on its way out of the caller's buffers (they may be re-used afterwards)
    bool sync();
    bool close();
    const std::string& error() const { return err_; }
    uint64_t bytes_written() const { return pos_; }

  private:
    bool put(const void* p, size_t n);
    bool flush_small();
    bool patch_size8(uint64_t at, uint64_t value);
    bool flush_cluster();
    bool pwrite_all(const void* p, size_t n, uint64_t at);
    bool write_piece(const uint8_t* p, size_t n, uint64_t at);
    bool use_mmap_ = false;               // output on tmpfs: large pieces go through mappings (see write_piece)
    uint64_t file_size_ = 0;              // size the file has been grown to (mmap path)
    void writer_loop();
    int fd_ = -1;
    std::string err_;
    uint64_t pos_ = 0, segment_data_start_ = 0;
    std::vector<uint8_t> small_;          // element headers not yet written
    uint64_t small_at_ = 0;               // file offset of small_[0]
    uint64_t cluster_size_at_ = 0, cluster_data_start_ = 0;   // file offsets of the open Cluster's size field / first child
    int64_t cluster_time_ = -1;
    uint64_t seekhead_at_ = 0, info_at_ = 0, tracks_at_ = 0, attachments_at_ = 0;
    struct Cue { int64_t time; int track; uint64_t cluster_pos; };
    std::vector<Cue> cues_;               // first video block of every Cluster
    bool cluster_has_cue_ = false;
    uint64_t cluster_pos_ = 0;            // file offset of the open Cluster's ID
    int video_track_ = 0;                 // first video track (0: none)
    // writer pool
    struct Job { const uint8_t* p; size_t n; uint64_t at; };
    std::deque<Job> jobs_;
    std::vector<std::thread> writers_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    size_t inflight_ = 0;
    bool stop_ = false, io_failed_ = false;
};

}  // namespace b200
