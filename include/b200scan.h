/* b200scan.h — C ABI of the analysis-pass helpers.
 *
 * Before RAWcooked launches the encoder, its input parsers touch every byte of every source file, one file at a time on
 * one core (Source/CLI/Main.cpp:280-292):
 *   --hash            input_base::Hash, Source/Lib/Utils/FileIO/Input_Base.cpp:54-81: MD5 (RFC 1321,
 *                     Source/Lib/ThirdParty/md5) of the whole file, kept in the reversibility data
 *   --check-padding   dpx::ParseBuffer, Source/Lib/Uncompressed/DPX/DPX.cpp:500-608: are the padding bits of the payload
 *                     zero? If not, the payload masked to its padding bits is kept ("In" data) so that decoding can put
 *                     them back
 * With the encode at hundreds of frames per second that pass (~0.6 GB/s = ~12 4K frames/s per core) is the next wall. The
 * entry points below do both on a B200: MD5 is serial inside a message, so one lane hashes one file and a batch of files
 * fills the warps; the padding scan is a plain streaming pass at HBM speed. No CPU fallback.
 *
 * Conventions as in b200enc.h (status codes, b200_last_error()).
 */
#ifndef B200SCAN_H
#define B200SCAN_H

#include <stddef.h>
#include <stdint.h>

#include "b200enc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_scan b200_scan;

/* max_items buffers per call, max_bytes in total per call (device staging of the host entry points). */
int b200_scan_open(int32_t device, int32_t max_items, size_t max_bytes, b200_scan** out);
void b200_scan_close(b200_scan* s);

/* MD5 of n byte ranges in HOST memory; digests receives 16 bytes per range. Synchronous, includes the H2D copy. */
int b200_md5_host(b200_scan* s, const uint8_t* const* data, const size_t* len, int32_t n, uint8_t* digests);
/* MD5 of n byte ranges of ONE device buffer (d_base + off[i], len[i]); asynchronous on `stream` up to the final copy of the
 * digests, which the call waits for. Ranges that start on a 16-byte boundary are read with 16-byte loads. */
int b200_md5_device(b200_scan* s, const void* d_base, const size_t* off, const size_t* len, int32_t n, uint8_t* digests, void* stream);

/* The padding-bit test of dpx::ParseBuffer (DPX.cpp:500-608) on n payloads (HOST memory, b200_ffv1_frame_bytes() each) of
 * one DPX layout (b200_layout 0..7). For payload i: nonzero[i] = tested units (bytes of a Filled sample, last 32-bit word
 * of a Packed row) with a padding bit set; first[i] = byte offset of the first of them in the payload (UINT64_MAX if none —
 * the reference's In_FirstNonZero); when masked != NULL and masked[i] != NULL, masked[i] receives the payload with
 * everything but the tested padding bits cleared (the reference's `In` buffer). */
int b200_padding_host(b200_scan* s, uint32_t width, uint32_t height, int32_t layout, const uint8_t* const* payloads, int32_t n,
                      uint64_t* nonzero, uint64_t* first, uint8_t* const* masked);
/* Same on payloads already in device memory (n payloads back to back); d_masked may be NULL. */
int b200_padding_device(b200_scan* s, uint32_t width, uint32_t height, int32_t layout, const void* d_payloads, int32_t n,
                        uint64_t* nonzero, uint64_t* first, void* d_masked, void* stream);

/* `-f framemd5` (the second output RAWcooked adds with --framemd5, Source/CLI/Output.cpp:312-332): MD5 of every frame as
 * FFmpeg's rawvideo encoder emits it, i.e. in the pix_fmt its dpx / tiff decoder produces for the flavor (8 bit: rgb24;
 * 10 / 12 bit: gbrp10le / gbrp12le, planes G, B, R; 16 bit: rgb48le / rgb48be by the file's byte order), rows without padding.
 * b200_rawvideo_bytes() is the `size` column of the framemd5 file, b200_rawvideo_pix_fmt() the pix_fmt name. */
size_t b200_rawvideo_bytes(uint32_t width, uint32_t height, int32_t layout);
const char* b200_rawvideo_pix_fmt(int32_t layout);
int b200_framemd5_host(b200_scan* s, uint32_t width, uint32_t height, int32_t layout, const uint8_t* const* payloads, int32_t n,
                       uint8_t* digests);
int b200_framemd5_device(b200_scan* s, uint32_t width, uint32_t height, int32_t layout, const void* d_payloads, int32_t n,
                         uint8_t* digests, void* stream);

/* Device time in microseconds of the last call's kernel ([0]) and bytes it read ([1]). */
int b200_scan_stats(const b200_scan* s, uint64_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif /* B200SCAN_H */
