/* b200dec.h — C ABI of the B200-native FFV1 decode + compare path (the `--check` side of RAWcooked).
 *
 * What this boundary replaces.  `rawcooked --check` (and `--all`) decodes every FFV1 packet of the
 * Matroska file on the CPU and compares the result with the source file, byte for byte:
 *   Source/Lib/CoDec/Wrapper.cpp:71-128          ffv1_wrapper::OutOfBand / ::Process (the base_wrapper
 *                                                interface, Source/Lib/CoDec/Wrapper.h:30-41)
 *   Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:134-228 packet -> slices (tail walk)
 *   Source/Lib/CoDec/FFV1/FFV1_Slice.cpp:210-472 slice header, LineThenPlane, Line (predict + context + Sample_Delta)
 *   Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:71-304  rangecoder::b / ::u / ::s
 *   Source/Lib/Transform/Transform.cpp:29-420    inverse RCT + DPX/TIFF byte layouts
 *   Source/Lib/Utils/FileIO/FileWriter.cpp:464-727  compare with the file on disk
 * SURVEY.md measures that path at 3.5 s per 4K frame and core; once the encode runs at hundreds of
 * frames per second it is all of `rawcooked --all`. The entry points below do the same work on a B200:
 * every slice of every frame of a batch is decoded by its own lane (the decoder is serial per slice in
 * both the model and the coder, so the parallelism is slices x frames), the rows go through the inverse
 * RCT and into the file's byte layout warp-wide, and are either written out or compared in place with
 * the source payloads.
 *
 * Supported streams: what b200enc.h's encoder and ffmpeg's `-c:v ffv1 -level 3 -coder 1 -g 1` produce for RGB
 * sources: version 3 (micro >= 4), range coder (coder_type 1 or 2), colorspace_type 1 (JPEG2000-RCT), no alpha,
 * intra-only, 8..16 bits, ec 0 or 1, up to 8 quantisation-table sets, initial states not coded. Anything else
 * is refused by b200_ffv1_dec_open with B200_ERR_INVALID. No CPU fallback.
 *
 * Conventions as in b200enc.h (status codes, b200_last_error()).
 */
#ifndef B200DEC_H
#define B200DEC_H

#include <stddef.h>
#include <stdint.h>

#include "b200enc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ffv1_dec_cfg {
    uint32_t width;
    uint32_t height;
    int32_t  layout;            /* b200_layout of the payload to produce / compare with; its bit depth must equal the stream's */
    int32_t  max_frames;        /* frames per call (>= 1) */
    int32_t  device;
    int32_t  slices_per_warp;   /* 0 = chosen per call from the slices in flight (or B200_DEC_SPW); 1..32: slices that share a warp, one lane each */
    int32_t  reserved[6];       /* must be 0 */
} b200_ffv1_dec_cfg;

typedef struct b200_ffv1_dec b200_ffv1_dec;

/* What parameters::Parse reads from the ConfigurationRecord (FFV1_Parameters.cpp:23-183). Host only, no device needed. */
typedef struct b200_ffv1_params {
    int32_t version, micro_version, coder_type /* as sent: 1 or 2 */, colorspace_type, bits_per_raw_sample;
    int32_t chroma_planes, log2_h_chroma_subsample, log2_v_chroma_subsample, alpha_plane;
    int32_t num_h_slices, num_v_slices, quant_table_set_count, ec, intra;
    int32_t context_count[8];   /* per quantisation-table set */
    int32_t crc_ok;             /* the record's CRC-32 parity is 0 (FFV1_Frame.cpp:114-117) */
} b200_ffv1_params;
int b200_ffv1_parse_config_record(const uint8_t* record, size_t record_len, b200_ffv1_params* out);

/* Per-frame status bits reported by the decode calls (0 = the frame decoded cleanly). */
enum {
    B200_DEC_BAD_TAIL     = 1,   /* slice sizes do not add up to the packet (FFV1_Frame.cpp:172-181) or wrong slice count */
    B200_DEC_BAD_CRC      = 2,   /* FFV1-SLICE-slice_crc_parity (FFV1_Slice.cpp:247-249) */
    B200_DEC_BAD_HEADER   = 4,   /* FFV1-SLICE-slice_xywh / quant_table_index (FFV1_Slice.cpp:113-177), not a keyframe */
    B200_DEC_UNDERRUN     = 8,   /* FFV1-SLICE-SliceContent (FFV1_Slice.cpp:345-346) */
    B200_DEC_JUNK         = 16,  /* FFV1-SLICE-JUNK (FFV1_Slice.cpp:296-297) */
    B200_DEC_ERROR_STATUS = 32   /* FFV1-SLICE-error_status (FFV1_Slice.cpp:307-309) */
};

/* Create a decoder for the stream described by `record` (Matroska CodecPrivate of the V_FFV1 track). */
int b200_ffv1_dec_open(const b200_ffv1_dec_cfg* cfg, const uint8_t* record, size_t record_len, b200_ffv1_dec** out);
void b200_ffv1_dec_close(b200_ffv1_dec* dec);

/* Decode n packets lying in HOST memory (packets[i], packet_len[i]) into payloads in HOST memory (frames[i], each
 * b200_ffv1_frame_bytes() long; row padding is written as zeros). status[i] receives the B200_DEC_* bits of frame i.
 * Synchronous, includes H2D / D2H. Returns 0 when the call itself worked (look at status[] for the streams). */
int b200_ffv1_decode_host(b200_ffv1_dec* dec, const uint8_t* const* packets, const size_t* packet_len, int32_t n,
                          uint8_t* const* frames, uint32_t* status);

/* The `--check` operation: decode n packets (HOST memory) and compare them on the GPU with the source payloads (HOST
 * memory, sources[i]). mismatch[i] = storage units of frame i that differ (bytes; 32-bit words for 12-bit packed DPX;
 * row padding is not pixel data and is not compared). Nothing is copied back but the two small arrays. */
int b200_ffv1_check_host(b200_ffv1_dec* dec, const uint8_t* const* packets, const size_t* packet_len, int32_t n,
                         const uint8_t* const* sources, uint64_t* mismatch, uint32_t* status);

/* Device-resident variant, asynchronous on `stream` (cudaStream_t, 0 = default): packets in ONE device buffer `d_packets`
 * at pkt_off[i] / pkt_len[i] (host arrays, e.g. what b200_ffv1_packets_device() reported for the encoder's arena);
 * `d_out` (n payloads back to back, may be NULL) receives the decoded payloads, `d_sources` (n payloads back to back, may
 * be NULL) is compared with them. Collect with b200_ffv1_dec_result() after the stream has been synchronised. */
int b200_ffv1_decode_device(b200_ffv1_dec* dec, const void* d_packets, const size_t* pkt_off, const size_t* pkt_len, int32_t n,
                            void* d_out, const void* d_sources, void* stream);
int b200_ffv1_dec_result(b200_ffv1_dec* dec, uint64_t* mismatch, uint32_t* status, int32_t n);

/* Device time of the kernels of the last decode call in microseconds: [0] k_dec_index, [1] k_decode, [2] total; and
 * [3] slices decoded, [4] samples decoded. Valid after b200_ffv1_dec_result() / the host calls. */
int b200_ffv1_dec_stats(const b200_ffv1_dec* dec, uint64_t stats[8]);

#ifdef __cplusplus
}
#endif
#endif /* B200DEC_H */
