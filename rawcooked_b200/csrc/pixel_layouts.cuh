// Pixel layouts of the source payloads (include/b200enc.h b200_layout): the three stored components of a pixel, exactly as
// the file holds them. Inverse of the byte layouts the reference writes in /root/reference/Source/Lib/Transform/Transform.cpp:70-420.
#pragma once
#include <cstdint>

#include "../../include/b200enc.h"

namespace b200 {

// pixel fetch: the three stored components of pixel x in a payload row (value bits only)
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

__device__ __forceinline__ void load_rgb(const uint8_t* __restrict__ row, int layout, int x, int& r, int& g, int& b) {
    switch (layout) {
        case B200_DPX_RGB_8: case B200_TIFF_RGB_8: {
            const uint8_t* q = row + 3 * x;
            r = q[0]; g = q[1]; b = q[2];
            break;
        }
        case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE: {
            uint32_t v = *reinterpret_cast<const uint32_t*>(row + 4 * x);
            if (layout == B200_DPX_RGB_10_FILLED_A_BE) v = bswap32(v);
            r = (v >> 22) & 1023; g = (v >> 12) & 1023; b = (v >> 2) & 1023;
            break;
        }
        case B200_DPX_RGB_12_FILLED_A_LE: case B200_DPX_RGB_16_LE: case B200_TIFF_RGB_16_LE: {
            const uint16_t* q = reinterpret_cast<const uint16_t*>(row + 6 * x);
            int sh = layout == B200_DPX_RGB_12_FILLED_A_LE ? 4 : 0;
            r = q[0] >> sh; g = q[1] >> sh; b = q[2] >> sh;
            break;
        }
        case B200_DPX_RGB_12_FILLED_A_BE: case B200_DPX_RGB_16_BE: case B200_TIFF_RGB_16_BE: {
            const uint16_t* q = reinterpret_cast<const uint16_t*>(row + 6 * x);
            int sh = layout == B200_DPX_RGB_12_FILLED_A_BE ? 4 : 0;
            uint32_t a = q[0], c = q[1], d = q[2];
            r = (int)(__byte_perm(a, 0, 0x4401)) >> sh;
            g = (int)(__byte_perm(c, 0, 0x4401)) >> sh;
            b = (int)(__byte_perm(d, 0, 0x4401)) >> sh;
            break;
        }
        case B200_DPX_RGB_12_PACKED_BE: {
            // component k = 3x + c lives at bit 12k (LSB first) of the row seen as big-endian 32-bit words
            const uint32_t* wds = reinterpret_cast<const uint32_t*>(row);
            int v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                uint32_t bit = (uint32_t)(3 * x + c) * 12u;
                uint32_t wi = bit >> 5, sh = bit & 31;
                uint32_t lo = bswap32(wds[wi]);
                uint32_t hi = sh > 20 ? bswap32(wds[wi + 1]) : 0u;   // 12 bits straddle only when sh > 20
                v[c] = (int)(__funnelshift_r(lo, hi, sh) & 0xFFFu);
            }
            r = v[0]; g = v[1]; b = v[2];
            break;
        }
        default: r = g = b = 0;
    }
}


}  // namespace b200
