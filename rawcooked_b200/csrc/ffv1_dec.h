// B200 FFV1 decoder (the `--check` side): stream description parsed from the ConfigurationRecord on the host, argument
// block and launchers of the decode kernels. Semantics = the reference decoder's:
//   record          /root/reference/Source/Lib/CoDec/FFV1/FFV1_Parameters.cpp:23-253
//   packet -> slices Source/Lib/CoDec/FFV1/FFV1_Frame.cpp:134-228
//   slice           Source/Lib/CoDec/FFV1/FFV1_Slice.cpp:113-177 (header), :210-318 (Parse), :406-472 (LineThenPlane, Line)
//   range decoder   Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:25-171
//   pixels          Source/Lib/Transform/Transform.cpp:29-420
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct Ffv1DecStream {
    int version = 0, micro = 0, coder_type_sent = 0, colorspace = 0, bits = 0;
    int chroma_planes = 0, log2_h = 0, log2_v = 0, alpha = 0;
    int num_h = 0, num_v = 0, nsets = 0, ec = 0, intra = 0;
    int nctx[8] = {0};
    bool states_coded = false;
    bool crc_ok = false;
    uint8_t one_state[256], zero_state[256];
    std::vector<int16_t> qtab;      // [nsets][5][256]
};
// 0, or a negative b200_status with text in *err
int parse_config_record(const uint8_t* rec, size_t n, Ffv1DecStream* out, std::string* err);

constexpr int kDecWarpsPerCta = 4;

struct DecArgs {
    int32_t W, H, layout, bits, bits_max, swap_bg, num_h, num_v, nslices, ec, tail, nsets, maxctx, spw;
    int32_t nctx[8];
    uint32_t row_bytes;
    size_t frame_bytes;
    const int16_t* qtab;          // [nsets][5][256]
    const uint8_t* trans;         // [512]: next state after a 0 ([0..255]) / after a 1 ([256..511])
    const uint32_t* crc_table;    // [256]
    // per batch
    const uint8_t* packets;
    const uint64_t* pkt_off;      // [n]
    const uint64_t* pkt_len;      // [n]
    uint64_t* sl_off;             // [n][nslices]  offset of slice k (in tail-walk order) inside the packet buffer
    uint32_t* sl_size;            // [n][nslices]  bytes including the tail; 0 = missing
    uint8_t* states;              // [n][nslices][2][maxctx*32]
    int32_t* lines;               // [n][nslices][3 planes][3 rows][wpad]
    int32_t wpad;
    uint8_t* out;                 // n payloads back to back, or null
    const uint8_t* cmp;           // n payloads back to back, or null
    unsigned long long* mismatch; // [n]
    uint32_t* status;             // [n]
    unsigned long long* counters; // [0] slices decoded, [1] samples decoded
};

cudaError_t launch_dec_index(const DecArgs& a, int nframes, cudaStream_t s);
cudaError_t launch_decode(const DecArgs& a, int nframes, cudaStream_t s);

}  // namespace b200
