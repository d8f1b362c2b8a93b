// Device-side interface of the B200 FLAC encoder (see flac_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200 {

constexpr int kFlacThreads = 256;

struct FlacArgs {
    const uint8_t* pcm;        // interleaved little-endian PCM exactly as in the WAV data chunk (8-bit unsigned, 16/24-bit signed)
    uint64_t n_samples;        // per channel, in this call
    uint64_t first_frame;      // frame number of block 0 of this call
    int32_t channels, bits, sample_rate, block_size;
    int32_t use_lpc;           // 1: LPC orders 1..8 compete with the fixed predictors (ffmpeg's level 5), 0: fixed only
    uint32_t* out;             // [blocks][frame_words], zeroed; ends up holding the frame bytes
    uint32_t frame_words;
    uint32_t* frame_len;       // [blocks] bytes
};

cudaError_t launch_flac(const FlacArgs& a, int nblocks, cudaStream_t s);

}  // namespace b200
