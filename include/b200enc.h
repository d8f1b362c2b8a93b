/* b200enc.h — C ABI of the B200-native FFV1 (+FLAC) encode path behind RAWcooked.
 *
 * What this boundary replaces.  RAWcooked has no in-process encoder: after analysing the
 * inputs it builds an `ffmpeg …` command line and runs it with system()
 * (reference: Source/CLI/Output.cpp:36-378, the system() call is Output.cpp:356; the binary
 * name comes from --bin-name, Source/CLI/Global.cpp:543-550, default "ffmpeg" :913-914).
 * The per-frame work ffmpeg does for that command (FFV1 v3 slice encode with RAWcooked's
 * option set `-coder 1 -context 1 -g 1 -level 3 -slicecrc 1 -slices N`,
 * Source/CLI/Global.cpp:938-989) is what the entry points below do on a B200.
 * The bitstream they emit is what the reference's own decoder consumes in `--check`
 * (Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:134-228, FFV1_Slice.cpp:210-318).
 *
 * Conventions: plain C, caller-owned buffers, no exceptions cross the boundary, every
 * function returns 0 on success or a negative b200_status; b200_last_error() gives text.
 * There is NO CPU fallback: without a CUDA device every encode entry point fails with
 * B200_ERR_NO_DEVICE.
 */
#ifndef B200ENC_H
#define B200ENC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b200_status {
    B200_OK = 0,
    B200_ERR_INVALID = -1,      /* bad argument / unsupported configuration            */
    B200_ERR_NO_DEVICE = -2,    /* no CUDA device (there is no CPU fallback)           */
    B200_ERR_CUDA = -3,         /* a CUDA runtime call or a kernel failed              */
    B200_ERR_OVERFLOW = -4,     /* an output or intermediate buffer was too small      */
    B200_ERR_IO = -5,           /* file open/read/write failed (CLI entry point)       */
    B200_ERR_EXISTS = -6        /* output exists and -n was given (CLI entry point)    */
} b200_status;

/* Source pixel layouts = the image payload of the file exactly as stored on disk.
 * Values 0..7 equal the reference's dpx::flavor indices
 * (Source/Lib/Uncompressed/DPX/DPX.h:38-46; byte layouts at Source/Lib/Transform/Transform.cpp:70-420),
 * 32..34 are tiff::flavor RGB flavors (Source/Lib/Uncompressed/TIFF/TIFF.h:36-40; Transform.cpp:1048-1055). */
typedef enum b200_layout {
    B200_DPX_RGB_8              = 0,  /* R,G,B bytes                                              */
    B200_DPX_RGB_10_FILLED_A_LE = 1,  /* 32-bit word R<<22|G<<12|B<<2, little endian              */
    B200_DPX_RGB_10_FILLED_A_BE = 2,  /* same word, big endian                                    */
    B200_DPX_RGB_12_FILLED_A_LE = 3,  /* 3 x 16-bit (v<<4), little endian                         */
    B200_DPX_RGB_12_PACKED_BE   = 4,  /* 12-bit samples LSB-first in big-endian 32-bit words,     */
                                      /* each row padded to a 32-bit boundary                     */
    B200_DPX_RGB_12_FILLED_A_BE = 5,  /* 3 x 16-bit (v<<4), big endian                            */
    B200_DPX_RGB_16_LE          = 6,  /* 3 x 16-bit little endian                                 */
    B200_DPX_RGB_16_BE          = 7,  /* 3 x 16-bit big endian                                    */
    B200_TIFF_RGB_8             = 32, /* R,G,B bytes, rows tightly packed                         */
    B200_TIFF_RGB_16_LE         = 33,
    B200_TIFF_RGB_16_BE         = 34
} b200_layout;

/* Encoder configuration = the subset of ffmpeg's ffv1 options RAWcooked emits
 * (Source/CLI/Global.cpp:938-989; slices: Source/CLI/Output.cpp:41-57). */
typedef struct b200_ffv1_cfg {
    uint32_t width;
    uint32_t height;
    int32_t  layout;      /* b200_layout; fixes bits_per_raw_sample (8/10/12/16)                 */
    int32_t  slices;      /* `-slices N`; 0 = pick like ffmpeg; grid = b200_ffv1_slice_grid()    */
    int32_t  context;     /* `-context`: 0 small, 1 large context model (RAWcooked default 1)     */
    int32_t  coder;       /* `-coder`: 1 (range coder; sent as coder_type 2 like ffmpeg) only     */
    int32_t  slicecrc;    /* `-slicecrc`: 1 (ec = 1, RAWcooked default) or 0                      */
    int32_t  max_frames;  /* frames per batch the handle must be able to hold in flight (>=1)     */
    int32_t  device;      /* CUDA device ordinal                                                  */
    int32_t  reserved[7]; /* must be 0 */
} b200_ffv1_cfg;

typedef struct b200_ffv1_enc b200_ffv1_enc;

/* `-slices N` -> (num_h_slices, num_v_slices) with ffmpeg's search (the set of
 * valid counts is pinned by the reference's test/slices.sh:12). Returns B200_ERR_INVALID if N has no grid. */
int b200_ffv1_slice_grid(uint32_t width, uint32_t height, int32_t slices, int32_t* num_h, int32_t* num_v);

/* Bytes of image payload one frame has in `layout` (what the reference computes at
 * Source/Lib/Uncompressed/DPX/DPX.cpp:460-483). 0 if unsupported. */
size_t b200_ffv1_frame_bytes(uint32_t width, uint32_t height, int32_t layout);

/* Create an encoder (allocates all device memory for max_frames frames in flight). */
int b200_ffv1_open(const b200_ffv1_cfg* cfg, b200_ffv1_enc** out);
void b200_ffv1_close(b200_ffv1_enc* enc);

/* FFV1 ConfigurationRecord = Matroska CodecPrivate of the V_FFV1 track
 * (what the reference parses at Source/Lib/CoDec/FFV1/FFV1_Parameters.cpp:23-183 after the CRC
 * check at FFV1_Frame.cpp:114-117). Returns its size; copies min(size, cap) bytes. */
size_t b200_ffv1_config_record(const b200_ffv1_enc* enc, uint8_t* out, size_t cap);

/* The same record computed on the host alone, without opening an encoder (no device needed): what a muxer that writes the
 * track header before the first frame is encoded would call. Returns the size (0 and b200_last_error() on a bad configuration);
 * copies min(size, cap) bytes. */
size_t b200_ffv1_config_record_for(const b200_ffv1_cfg* cfg, uint8_t* out, size_t cap);

/* Upper bound of one packet's size, for sizing `out` below. */
size_t b200_ffv1_max_packet_bytes(const b200_ffv1_enc* enc);

/* Encode `n_frames` (<= max_frames) frames whose payloads lie in HOST memory at frames[i]
 * (each b200_ffv1_frame_bytes() long). Packets are written back to back into `out`
 * (host memory, capacity `out_cap`); packet i occupies [out_off[i], out_off[i] + out_len[i]).
 * One packet = one Matroska SimpleBlock payload (one keyframe, all slices, slice CRCs).
 * The call includes host->device and device->host copies and is synchronous. */
int b200_ffv1_encode_host(b200_ffv1_enc* enc, const uint8_t* const* frames, int32_t n_frames,
                          uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len);

/* Asynchronous half of b200_ffv1_encode_host: enqueues the band-by-band host->device copies and the kernels and returns at
 * once; the frames must stay valid (and should be pinned) until the batch has been collected. Up to TWO batches may be in
 * flight: b200_ffv1_fetch_packets() / b200_ffv1_packets_device() always collect the OLDEST one, so
 *     submit(0); submit(1); fetch -> 0; submit(2); fetch -> 1; ...
 * moves the packets of batch i over PCIe while batch i+1 is being coded (two result sets on the device; the second one is
 * allocated when first needed). A third submit without a fetch gives up the oldest batch. Do not mix with
 * b200_ffv1_encode_device while host batches are in flight. */
int b200_ffv1_submit_host(b200_ffv1_enc* enc, const uint8_t* const* frames, int32_t n_frames);
/* Optional, ahead of a submit: starts the host->device copy of the NEXT batch (whole frames, into the staging buffer the batch
 * in flight does not use) and returns at once. Called before the fetch of the previous batch,
 *     submit(0); submit(1); prefetch(2); fetch -> 0; submit(2); prefetch(3); fetch -> 1; submit(3); ...
 * the upload of batch i+1 and the download of batch i-1 share the PCIe link in both directions while batch i is coded; the
 * following b200_ffv1_submit_host with the same frames then only launches kernels. Without a batch in flight it does nothing. */
int b200_ffv1_prefetch_host(b200_ffv1_enc* enc, const uint8_t* const* frames, int32_t n_frames);

/* Device-resident variant: `d_frames` is ONE device buffer holding n_frames payloads back to back
 * (stride b200_ffv1_frame_bytes()); packets are produced into an internal device buffer.
 * `stream` is a cudaStream_t (0 = default stream); the call is asynchronous on it.
 * After a stream sync, b200_ffv1_packets_device() returns the device pointer of the packet arena and
 * copies the per-frame offsets/lengths to the host arrays; b200_ffv1_fetch_packets() copies the bytes. */
int b200_ffv1_encode_device(b200_ffv1_enc* enc, const void* d_frames, int32_t n_frames, void* stream);
int b200_ffv1_packets_device(b200_ffv1_enc* enc, const void** d_arena, size_t* out_off, size_t* out_len, int32_t n_frames);
int b200_ffv1_fetch_packets(b200_ffv1_enc* enc, uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len, int32_t n_frames);
/* Waits for the OLDEST batch in flight and reports its packet layout (offsets / lengths in the arena, total bytes) without
 * collecting it: the caller can size (and pin) its output buffer before b200_ffv1_fetch_packets(). */
int b200_ffv1_packet_sizes(b200_ffv1_enc* enc, size_t* out_off, size_t* out_len, int32_t n_frames, size_t* total_bytes);

/* Counters of the last encode call: [0] kernels launched, [1] range-coder bins coded,
 * [2] samples coded, [3] bytes of packets produced; plus, when timing is enabled with b200_ffv1_set_timing()
 * (the kernels then run one after the other on the caller's stream, each bracketed by CUDA events), the summed
 * device time in microseconds of [4] k_model, [5] k_range, [6] k_scan+k_pack, [7] k_emit over all bands.
 * Valid after b200_ffv1_packets_device() / b200_ffv1_fetch_packets(). */
int b200_ffv1_stats(const b200_ffv1_enc* enc, uint64_t stats[8]);

/* Geometry of the encoder: [0] num_h_slices, [1] num_v_slices, [2] bands per frame (= launches of each band
 * kernel per encode call), [3] rows per band, [4] slices per frame, [5] widest slice, [6] tallest slice,
 * [7] bits_per_raw_sample. */
int b200_ffv1_info(const b200_ffv1_enc* enc, int32_t info[8]);
int b200_ffv1_set_timing(b200_ffv1_enc* enc, int32_t enabled);

/* ---- FLAC (what `-c:a flac` asks of the child process, /root/reference/Source/CLI/Global.cpp:949-950). One frame per block,
 * fixed block size, independent channels, fixed predictors and LPC (orders 1..8, 15-bit coefficients: ffmpeg's level 5) +
 * partitioned Rice, all on the GPU. The frames are what the
 * reference's vendored libFLAC decodes in `--check` (Source/Lib/CoDec/Wrapper.cpp:138-219). */
typedef struct b200_flac_cfg {
    uint32_t sample_rate;
    uint32_t channels;    /* 1..8 */
    uint32_t bits;        /* 8 (unsigned WAV bytes), 16, 24 (signed little endian) */
    int32_t  block_size;  /* 0 = like ffmpeg: largest standard size <= 105 ms (4608 @ 48 kHz, 8192 @ 96 kHz); <= 16384 */
    int32_t  max_blocks;  /* blocks per encode call (0 = 256) */
    int32_t  device;
    int32_t  fixed_only;  /* 1 = fixed predictors only (ffmpeg's -compression_level 0..2); 0 = LPC competes too (the default) */
    int32_t  reserved[5]; /* must be 0 */
} b200_flac_cfg;
typedef struct b200_flac_enc b200_flac_enc;
int b200_flac_open(const b200_flac_cfg* cfg, b200_flac_enc** out);
void b200_flac_close(b200_flac_enc* enc);
int32_t b200_flac_block_size(const b200_flac_enc* enc);
size_t b200_flac_max_frame_bytes(const b200_flac_enc* enc);
/* Matroska CodecPrivate of the A_FLAC track: "fLaC" + STREAMINFO block (42 bytes); call after the last encode so that the
 * min/max frame sizes are known. The MD5 signature of the unencoded audio (which the reference's libFLAC verifies at the end
 * of the stream, Source/Lib/CoDec/Wrapper.cpp:189) is filled in when exactly total_samples samples went through this handle's
 * encode calls, in order; otherwise it is left 0 (= not computed, the decoder then skips the check). */
size_t b200_flac_codec_private(const b200_flac_enc* enc, uint64_t total_samples, uint8_t* out, size_t cap);
/* Encode n_samples (per channel) of interleaved PCM given exactly as in the WAV data chunk (host memory). Frames are written
 * back to back into `out`; frame i = block i, numbered first_frame + i. Synchronous, includes H2D/D2H. */
int b200_flac_encode_host(b200_flac_enc* enc, const uint8_t* pcm, uint64_t n_samples, uint64_t first_frame,
                          uint8_t* out, size_t out_cap, size_t* out_off, size_t* out_len, int32_t* n_frames);

/* ---- argv-compatible entry point: the `ffmpeg` command line RAWcooked builds (Source/CLI/Output.cpp:36-378). Returns the
 * process exit status (0 = success). `rawcooked -b b200enc` runs the executable that wraps this call. */
int b200enc_main(int argc, char** argv);
/* The b200enc executable ends with the job: with on != 0, b200enc_main skips the teardown of device and pinned host memory once
 * the output file is complete (seconds for tens of GB) and leaves it to process exit. Library callers keep the default (0). */
void b200enc_set_fast_exit(int on);

/* Text of the last error on this thread ("" if none). */
const char* b200_last_error(void);

/* Library/ABI version: (major<<16)|(minor<<8)|patch. */
uint32_t b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200ENC_H */
