// Device-side interface of the B200 FFV1 encoder: argument block and launchers of the band kernels.
//
//   k_model  (K1+K2+K3)  one CTA per (frame, slice, plane-set): unpack + forward RCT, median prediction, context
//                         quantisation, binarisation, and resolution of the adaptive probability state every bin sees.
//                         Emits, in bitstream order, one record per range-coder bin: a byte q = sp - 1 (sp = 8-bit
//                         probability of the coded value; q = 255 is a no-op used for padding) and one bit (the coded value)
//                         in a separate bit-plane.
//   k_range  (K4a)        one lane per (frame, slice): the strictly serial recurrence of `range` and the byte count;
//                         leaves a checkpoint every 64 records.
//   k_emit   (K4b)        one thread per 64 records: replays them from the checkpoint and adds their bytes into the
//                         slice stream.
//   k_scan / k_pack       slice sizes -> packet layout; compaction of the per-slice byte streams into packets, CRC, footer.
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
#define B200_MODEL_THREADS 768
#endif
constexpr int kModelThreads = B200_MODEL_THREADS;   // one CTA per SM when the large context model (162 KB) is in smem
constexpr int kModelWarps = kModelThreads / 32;
constexpr int kMaxHeaderBins = 96;
constexpr int kBlockRecs = 128;                     // every plane-row segment is padded to a multiple of this many records
constexpr int kMaxSeg = 4;                          // column segments a plane-row may be coded in (stage capacity)
constexpr int kModelSmemReserve = 4 * 1024;         // left free per SM so that k_range CTAs (no shared memory) co-reside

// persistent range-coder registers of one slice between bands
struct CoderState {
    uint32_t range;
    uint32_t pos;           // bytes produced so far
};

struct EncArgs {
    // stream description
    int32_t W, H, layout, bits, sbits, swap_bg, nslices, nctx, is5, ec;
    int32_t sstride;              // state bytes per context in the model kernel: 32, or 28 (27 used) for 8-bit streams
    uint32_t row_bytes;
    size_t frame_bytes;
    int32_t band_rows, nbands, wmax, hmax;
    int32_t stage_cap;            // records of one plane-row segment k_model stages in shared memory
    int32_t nseg;                 // column segments per plane-row reserved in rowcnt (1..kMaxSeg)
    const SliceGeom* geom;        // [nslices]
    const int16_t* qtab;          // [5][256]
    const uint8_t* t1q;           // [256] t1q[q] = one_state[q + 1]
    const uint8_t* tpow;          // [5][256] tpow[i][s] = one_state applied 2^i times to state s
    const uint16_t* hdr_bins;     // [nslices][kMaxHeaderBins]  records (q | bit << 8)
    const int32_t* hdr_cnt;       // [nslices]
    const uint32_t* crc_table;    // [256]
    // per batch
    const uint8_t* in;            // frames back to back, stride frame_bytes
    uint8_t* state_save;          // [frames][nslices][2][align16(nctx*sstride)]
    uint8_t* qY;                  // [frames][nslices][capY]      q bytes of the Y stream
    uint8_t* qC;                  // [frames][nslices][capC]      q bytes of the Cb/Cr stream
    uint8_t* bY;                  // [frames][nslices][capY/8]    bit-plane
    uint8_t* bC;                  // [frames][nslices][capC/8]
    size_t capY, capC;            // records per (frame, slice) region, multiples of kBlockRecs
    uint32_t* rowcnt;             // [frames][nslices][band_rows][3][nseg]  records of each segment of the Y, Cb, Cr row
    uint2* ckptY;                 // [frames][nslices][capY/64]  (range, bytes so far) every 64 records
    uint2* ckptC;                 // [frames][nslices][capC/64]
    uint32_t* used;               // [frames][nslices][2]  64-record pieces in use in the Y / C stream of this band
    CoderState* cstate;           // [frames][nslices]
    uint8_t* scratch;             // [frames][nslices][slice_cap]
    size_t slice_cap;
    uint32_t* slice_size;         // [frames][nslices]  payload bytes (footer excluded)
    uint64_t* slice_off;          // [frames][nslices]  offset in arena
    uint64_t* frame_off;          // [frames] offset of packet in arena
    uint64_t* frame_len;          // [frames]
    uint8_t* arena;
    size_t arena_cap;
    uint32_t* work_ctr;           // [nbands] next work item of k_model's persistent grid (zeroed per encode call)
    int32_t model_ctas;           // CTAs of k_model's grid (the SMs of its partition, or SMs minus the ones left to k_range / k_emit)
    int32_t range_smem;           // dynamic shared memory k_range asks for (unused by the kernel): keeps it off k_model's SMs
                                  // when the device is not partitioned, keeps it at one CTA per SM inside a partition
    int32_t range_sms;            // SMs of k_range's partition (0: the device is not partitioned)
    uint32_t* flags;              // [0] overflow flag, [2..3] total bins, [16..] phase cycles (-DB200_PHASE_TIMING)
};

size_t model_smem_fixed(int nctx, int sstride, int wmax, int planes);
size_t model_smem_bytes(int nctx, int sstride, int wmax, int planes, int stage_cap);
cudaError_t launch_model(const EncArgs& a, int band, int nframes, cudaStream_t s);
cudaError_t launch_range(const EncArgs& a, int band, int nframes, cudaStream_t s);
cudaError_t launch_emit(const EncArgs& a, int nframes, cudaStream_t s);
cudaError_t launch_pack(const EncArgs& a, int nframes, cudaStream_t s);
cudaError_t configure_kernels(const EncArgs& a);

}  // namespace b200
