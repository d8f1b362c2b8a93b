// Host-side FFV1 v3 stream description for the B200 encoder: everything that is computed once per
// stream on the CPU (slice grid, quantisation tables, state-transition tables, ConfigurationRecord,
// per-slice header bins). Per-frame work is all on the GPU (ffv1_kernels.cu).
//
// Bitstream semantics follow what the reference's decoder reads:
//   ConfigurationRecord   /root/reference/Source/Lib/CoDec/FFV1/FFV1_Parameters.cpp:23-183 (+ :222-253 tables)
//   slice header          Source/Lib/CoDec/FFV1/FFV1_Slice.cpp:113-177
//   range coder           Source/Lib/CoDec/FFV1/FFV1_RangeCoder.cpp:71-171
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200 {

struct SliceGeom { int32_t x0, y0, w, h; };

struct Ffv1Stream {
    uint32_t width = 0, height = 0;
    int layout = 0;
    int bits = 0;          // bits_per_raw_sample
    int sbits = 0;         // bits of a coded sample (bits+1 after RCT; 9 for 8-bit)
    int swap_bg = 0;       // 9..15-bit RGB without alpha: b and g swapped (Transform.cpp:104,126,233,338,363)
    int num_h = 0, num_v = 0;
    int context = 1;       // quant table set used by every plane
    int ec = 1;
    int nctx = 0;          // contexts of the selected set
    int is5 = 0;           // 5-input context (tables 3,4 non-zero; FFV1_Slice.cpp:453)
    size_t row_bytes = 0, frame_bytes = 0;
    int16_t qtab[5][256];  // selected set
    uint8_t one_state[256], zero_state[256];
    std::vector<SliceGeom> slices;               // raster order (sy major)
    std::vector<uint8_t> config_record;
    // bins (state | bit << 8) that precede the first sample of each slice: [keyframe bin for slice 0] + header symbols
    std::vector<std::vector<uint16_t>> header_bins;
};

int layout_bits(int layout);
size_t layout_row_bytes(uint32_t width, int layout);
int slice_grid(uint32_t width, uint32_t height, int slices, int bits, int* num_h, int* num_v);
// returns 0 or a negative b200_status; fills err
int build_stream(uint32_t width, uint32_t height, int layout, int slices, int context, int ec, Ffv1Stream* out, const char** err);

void default_one_state(uint8_t one[256]);   // default state transitions (FFV1_Frame.cpp:35-55)
uint32_t crc32_mpeg(const uint8_t* d, size_t n, uint32_t crc = 0);
const uint32_t* crc32_mpeg_table();

}  // namespace b200
