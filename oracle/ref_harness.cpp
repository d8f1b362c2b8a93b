// TEST INFRASTRUCTURE ONLY (oracle/): packet-level harness around the UNMODIFIED
// reference FFV1 decoder of MediaArea/RAWcooked. Compiled by oracle/build_ref.sh
// together with the reference's own sources where they lie under /root/reference
// (nothing is copied), output goes to oracle/_ref/libref_ffv1dec.so.
//
// It drives exactly what `rawcooked --check` drives for one video packet:
//   ffv1_frame::OutOfBand  (Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:105-131)
//   ffv1_frame::Process    (Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:134-228)
//   -> slice::Parse -> Transform->From  (DPX/TIFF byte layout, Transform.cpp)
// and hands back the bytes of raw_frame plane 0, i.e. the image payload of the
// DPX/TIFF file the packet must reproduce.
#include "Lib/CoDec/FFV1/FFV1_Frame.h"
#include "Lib/Utils/RawFrame/RawFrame.h"
#include "Lib/Utils/CRC32/ZenCRC32.h"
#include "ThreadPool.h"
#include <cstring>
#include <cstdio>

extern "C" {

// container: 0 = DPX (flavor = dpx::flavor index, DPX.h:38-61), 1 = TIFF (tiff::flavor, TIFF.h:36-45)
// returns 0 on success, 1 if the reference decoder flagged an error (message in err), 2 on size problems
int ref_ffv1_decode(const uint8_t* extradata, size_t extradata_size,
                    uint32_t width, uint32_t height, int container, int flavor,
                    const uint8_t* pkt, size_t pkt_size,
                    uint8_t* out, size_t out_cap, size_t* out_size,
                    char* err, size_t err_cap, int threads)
{
    if (err && err_cap) err[0] = 0;
    ThreadPool* pool = nullptr;
    if (threads > 1) { pool = new ThreadPool(threads); pool->init(); }
    int rc = 0;
    {
        ffv1_frame F(pool);
        raw_frame R;
        R.Flavor = container == 0 ? raw_frame::flavor::DPX : raw_frame::flavor::TIFF;
        R.Flavor_Private = (uint64_t)flavor;
        F.RawFrame = &R;
        F.SetWidth(width);
        F.SetHeight(height);
        F.OutOfBand(extradata, extradata_size);
        if (F.ErrorMessage()) {
            if (err) snprintf(err, err_cap, "%s", F.ErrorMessage());
            rc = 1;
        } else {
            bool bad = F.Process(pkt, pkt_size);
            if (bad || F.ErrorMessage()) {
                if (err) snprintf(err, err_cap, "%s", F.ErrorMessage() ? F.ErrorMessage() : "Process returned true");
                rc = 1;
            }
            if (R.Planes().empty()) { rc = rc ? rc : 2; }
            else {
                const buffer& B = R.Plane(0)->Buffer();
                if (out_size) *out_size = B.Size();
                if (B.Size() > out_cap) rc = rc ? rc : 2;
                else memcpy(out, B.Data(), B.Size());
            }
        }
    }
    if (pool) { pool->shutdown(); delete pool; }
    return rc;
}

// the reference's default state-transition table (Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:35-55)
extern const state_transitions_struct default_state_transitions;
void ref_default_state_transitions(uint8_t out[256]) { memcpy(out, default_state_transitions.States, 256); }

// CRC-32 as the reference computes it (Source/Lib/Utils/CRC32/ZenCRC32.cpp:1097-1135)
uint32_t ref_crc32(const uint8_t* data, size_t size) { return ZenCRC32(data, size); }

}

// ---------------------------------------------------------------------------------------------------------------------
// FLAC: the reference's vendored libFLAC 1.3.2 decoder (Source/Lib/ThirdParty/flac), driven like flac_wrapper does
// (Source/Lib/CoDec/Wrapper.cpp:138-219: CodecPrivate first, then one frame per SimpleBlock).
#include "FLAC/stream_decoder.h"
namespace {
struct flac_mem {
    const uint8_t* data; size_t size, pos;
    int32_t* out; size_t cap, n;      // interleaved samples
    unsigned channels, bps; int error;
};
FLAC__StreamDecoderReadStatus flac_read(const FLAC__StreamDecoder*, FLAC__byte buffer[], size_t* bytes, void* client) {
    flac_mem* m = (flac_mem*)client;
    size_t left = m->size - m->pos;
    if (!left) { *bytes = 0; return FLAC__STREAM_DECODER_READ_STATUS_END_OF_STREAM; }
    if (*bytes > left) *bytes = left;
    memcpy(buffer, m->data + m->pos, *bytes);
    m->pos += *bytes;
    return FLAC__STREAM_DECODER_READ_STATUS_CONTINUE;
}
FLAC__StreamDecoderWriteStatus flac_write(const FLAC__StreamDecoder*, const FLAC__Frame* frame, const FLAC__int32* const buffer[], void* client) {
    flac_mem* m = (flac_mem*)client;
    m->channels = frame->header.channels; m->bps = frame->header.bits_per_sample;
    for (unsigned i = 0; i < frame->header.blocksize; i++)
        for (unsigned c = 0; c < frame->header.channels; c++) {
            if (m->n < m->cap) m->out[m->n] = buffer[c][i];
            m->n++;
        }
    return FLAC__STREAM_DECODER_WRITE_STATUS_CONTINUE;
}
void flac_error(const FLAC__StreamDecoder*, FLAC__StreamDecoderErrorStatus, void* client) { ((flac_mem*)client)->error++; }
}  // namespace

extern "C" {
// stream = CodecPrivate ("fLaC" + STREAMINFO) followed by the frames. Returns 0 ok; out_n = samples decoded (all channels).
int ref_flac_decode(const uint8_t* stream, size_t size, int32_t* out, size_t cap, size_t* out_n, unsigned* channels, unsigned* bps)
{
    flac_mem m = {stream, size, 0, out, cap, 0, 0, 0, 0};
    FLAC__StreamDecoder* d = FLAC__stream_decoder_new();
    if (!d) return 3;
    FLAC__stream_decoder_set_md5_checking(d, true);
    if (FLAC__stream_decoder_init_stream(d, flac_read, 0, 0, 0, 0, flac_write, 0, flac_error, &m) != FLAC__STREAM_DECODER_INIT_STATUS_OK) { FLAC__stream_decoder_delete(d); return 3; }
    FLAC__bool ok = FLAC__stream_decoder_process_until_end_of_stream(d);
    // finish() returns false when MD5 checking is on and the signature in STREAMINFO does not match the decoded audio
    // (a zero signature switches the check off): what flac_wrapper relies on, Source/Lib/CoDec/Wrapper.cpp:189
    const FLAC__bool md5_ok = FLAC__stream_decoder_finish(d);
    FLAC__stream_decoder_delete(d);
    if (ok && !m.error && !md5_ok) { if (out_n) *out_n = m.n; return 4; }
    if (out_n) *out_n = m.n;
    if (channels) *channels = m.channels;
    if (bps) *bps = m.bps;
    return (!ok || m.error) ? 1 : (m.n > m.cap ? 2 : 0);
}
}

// ---------------------------------------------------------------------------------------------------------------------
// MD5 as input_base::Hash computes it for --hash (Source/Lib/Utils/FileIO/Input_Base.cpp:54-81) with the reference's
// vendored md5.c (Source/Lib/ThirdParty/md5)
extern "C" {
#include "md5.h"
void ref_md5(const uint8_t* data, size_t size, uint8_t out[16])
{
    MD5_CTX c;
    MD5_Init(&c);
    size_t off = 0;
    while (off < size) {
        unsigned long n = size - off > 0x40000000ul ? 0x40000000ul : (unsigned long)(size - off);
        MD5_Update(&c, data + off, n);
        off += n;
    }
    MD5_Final(out, &c);
}
}
