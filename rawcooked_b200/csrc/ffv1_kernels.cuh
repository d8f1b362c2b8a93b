// Device-side interface of the B200 FFV1 encoder: argument blocks and launchers of the four kernels.
//
//   k_model  (K1+K2+K3)  one CTA per (frame, slice, plane-set): unpack + forward RCT, median prediction, context
//                         quantisation, binarisation, and resolution of the adaptive probability state every bin sees.
//                         Emits, in bitstream order, one 16-bit (state | bit << 8) record per range-coder bin.
//   k_code   (K4)         one lane per (frame, slice): the strictly serial part — range/low update, renormalisation,
//                         carry handling, byte output, running CRC-32, slice footer.
//   k_scan / k_pack       slice sizes -> packet layout; compaction of the per-slice byte streams into packets.
//
// Frames are processed in horizontal bands of `band_rows` rows so that the bin records (the only large intermediate)
// never exceed two band buffers regardless of how many frames are in flight.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "ffv1_host.h"

namespace b200 {

#ifndef B200_MODEL_THREADS
#define B200_MODEL_THREADS 512
#endif
constexpr int kModelThreads = B200_MODEL_THREADS;   // one CTA per SM when the large context model (162 KB) is in smem
constexpr int kModelWarps = kModelThreads / 32;
constexpr int kMaxHeaderBins = 96;

// persistent range-coder registers of one slice between bands
struct CoderState {
    uint32_t low, range;
    int32_t pending;        // outstanding byte, -1 = none yet
    uint32_t run;           // outstanding 0xFF count
    uint32_t pos;           // bytes written to the slice scratch so far
    uint32_t crc;           // running CRC-32 of those bytes
    uint32_t offY, offC;    // (unused between bands; kept for 32-byte alignment)
};

struct EncArgs {
    // stream description
    int32_t W, H, layout, bits, sbits, swap_bg, nslices, nctx, is5, ec;
    int32_t sstride;              // state bytes per context in the model kernel: 32, or 27 for 8-bit streams
    uint32_t row_bytes;
    size_t frame_bytes;
    int32_t band_rows, nbands, wmax, hmax;
    int32_t stage_cap;            // records of one plane-row k_model can stage in shared memory (the rest go straight to global)
    const SliceGeom* geom;        // [nslices]
    const int16_t* qtab;          // [5][256]
    const uint8_t* trans;         // [0..255] zero_state, [256..511] one_state
    const uint16_t* hdr_bins;     // [nslices][kMaxHeaderBins]
    const int32_t* hdr_cnt;       // [nslices]
    const uint32_t* crc_table;    // [256]
    // per batch
    const uint8_t* in;            // frames back to back, stride frame_bytes
    uint8_t* state_save;          // [frames][nslices][2][align16(nctx*sstride)]
    uint16_t* binsY;              // [frames][nslices][capY]
    uint16_t* binsC;              // [frames][nslices][capC]
    size_t capY, capC;            // elements per (frame, slice) region
    uint32_t* rowcnt;             // [frames][nslices][band_rows][3]  bins of the Y, Cb, Cr row
    uint2* ckptY;                 // [frames][nslices][capY/64]  (range, bytes so far) at the head of every 64-record block
    uint2* ckptC;                 // [frames][nslices][capC/64]
    uint32_t* used;               // [frames][nslices][2]  blocks in use in the Y / C stream of this band
    CoderState* cstate;           // [frames][nslices]
    uint8_t* scratch;             // [frames][nslices][slice_cap]
    size_t slice_cap;
    uint32_t* slice_size;         // [frames][nslices]  payload bytes (footer excluded)
    uint64_t* slice_off;          // [frames][nslices]  offset in arena
    uint64_t* frame_off;          // [frames] offset of packet in arena
    uint64_t* frame_len;          // [frames]
    uint8_t* arena;
    size_t arena_cap;
    uint32_t* flags;              // [0] overflow flag, [1] total bins lo, [2] total bins hi
};

size_t model_smem_fixed(int nctx, int sstride, int wmax, int planes);
size_t model_smem_bytes(int nctx, int sstride, int wmax, int planes, int stage_cap);
cudaError_t launch_model(const EncArgs& a, int band, int nframes, cudaStream_t s);
cudaError_t launch_range(const EncArgs& a, int band, int nframes, cudaStream_t s);
cudaError_t launch_emit(const EncArgs& a, int nframes, cudaStream_t s);
cudaError_t launch_pack(const EncArgs& a, int nframes, cudaStream_t s);
cudaError_t configure_kernels(int nctx, int sstride, int wmax, int stage_cap);

}  // namespace b200
