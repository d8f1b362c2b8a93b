#!/usr/bin/env python
"""Generates tests/golden/ffv1_golden.npz: FFmpeg's own bitstream for small seeded inputs.

RAWcooked's encode path is ffmpeg's `ffv1` encoder with the option set of
/root/reference/Source/CLI/Global.cpp:938-989. That encoder is a third-party dependency absent from the
reference tree; the instance pinned here is libavcodec 62.11.100 bundled in this image (driven by
oracle/avcodec_ffv1.py). For each case we store the input payload, libavcodec's ConfigurationRecord
(extradata) and its packet; tests/test_oracle.py requires oracle/ffv1_oracle.c to reproduce both
byte-for-byte, and the GPU tests require the CUDA path to do the same.

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import avcodec_ffv1 as A  # noqa: E402
from rawcooked_b200 import synth as S  # noqa: E402

# (width, height, layout, slices, context, kind, seed[, slicecrc])
CASES = [
    (64, 48, S.DPX_RGB_8, 4, 1, "grain", 11),
    (64, 48, S.DPX_RGB_8, 4, 0, "grain", 12),
    (64, 48, S.DPX_RGB_10_FA_LE, 4, 1, "grain", 13),
    (64, 48, S.DPX_RGB_10_FA_BE, 4, 1, "grain", 14),
    (64, 48, S.DPX_RGB_10_FA_BE, 4, 0, "flat", 15),
    (64, 48, S.DPX_RGB_12_FA_LE, 4, 1, "grain", 16),
    (64, 48, S.DPX_RGB_12_FA_BE, 4, 1, "grain", 17),
    (96, 50, S.DPX_RGB_12_PACKED_BE, 4, 1, "grain", 18),
    (100, 50, S.DPX_RGB_12_PACKED_BE, 6, 1, "grain", 19),   # slice edges not on 8-pixel blocks
    (64, 48, S.DPX_RGB_16_LE, 4, 1, "grain", 20),
    (64, 48, S.DPX_RGB_16_BE, 4, 1, "grain", 21),
    (64, 48, S.DPX_RGB_16_BE, 4, 0, "grain", 22),
    (100, 75, S.DPX_RGB_16_BE, 6, 1, "grain", 23),           # ragged: 100/3, 75/2
    (101, 77, S.DPX_RGB_10_FA_BE, 9, 1, "grain", 24),        # ragged 3x3
    (120, 90, S.DPX_RGB_16_BE, 24, 1, "white", 25),          # 6x4 grid, worst-case content
    (64, 48, S.DPX_RGB_16_BE, 4, 1, "zero", 26),
    (64, 48, S.DPX_RGB_16_BE, 4, 1, "const", 27),
    (64, 48, S.TIFF_RGB_8, 4, 1, "grain", 28),
    (64, 48, S.TIFF_RGB_16_LE, 4, 1, "grain", 29),
    (64, 48, S.TIFF_RGB_16_BE, 4, 1, "flat", 30),
    (640, 480, S.DPX_RGB_8, 16, 1, "grain", 1000),           # BASELINE config 1 frame (RAWcooked default 16 slices)
    (64, 48, S.DPX_RGB_16_BE, 4, 1, "grain", 31, 0),         # -slicecrc 0: ec = 0, 3-byte slice tail
    (100, 75, S.DPX_RGB_10_FA_BE, 6, 0, "flat", 32, 0),
    (64, 48, S.DPX_RGB_8, 4, 1, "grain", 33, 0),
    (67, 45, S.DPX_RGB_16_BE, 4, 1, "grain", 34),            # odd width: 16-bit DPX rows carry 2 bytes of line padding
    (67, 45, S.DPX_RGB_16_LE, 6, 1, "flat", 35),
    (65, 47, S.DPX_RGB_8, 4, 1, "grain", 36),                # 8-bit DPX rows padded to 32 bits
    (67, 45, S.TIFF_RGB_16_LE, 4, 1, "grain", 37),           # TIFF strips stay tight
]


def av_encode(w, h, layout, slices, context, R, G, B, slicecrc=1):
    bits = S.LAYOUT_BITS[layout]
    if bits == 8:
        fmt = "bgr0"
        planes = [np.stack([B, G, R, np.zeros_like(R)], -1).astype(np.uint8).reshape(h, 4 * w)]
    else:
        fmt = "gbrp%dle" % bits
        planes = [G, B, R]
    e = A.FFV1Encoder(w, h, fmt, slices, context=context, slicecrc=slicecrc)
    pkt = e.encode_planes(planes)
    rec = e.extradata
    e.close()
    return rec, pkt


def main():
    out = {"libavcodec": np.frombuffer(A.version().encode(), np.uint8)}
    meta = []
    for i, case in enumerate(CASES):
        w, h, layout, slices, context, kind, seed = case[:7]
        slicecrc = case[7] if len(case) > 7 else 1
        R, G, B = S.rgb_content(w, h, S.LAYOUT_BITS[layout], seed, kind)
        payload = S.pack_payload(R, G, B, layout)
        rec, pkt = av_encode(w, h, layout, slices, context, R, G, B, slicecrc)
        meta.append((w, h, layout, slices, context, seed, slicecrc))
        small = w * h <= 16384
        out["payload_%d" % i] = payload if small else np.zeros(0, np.uint8)   # big cases are regenerated from the seed
        out["record_%d" % i] = np.frombuffer(rec, np.uint8)
        out["packet_%d" % i] = np.frombuffer(pkt, np.uint8)
        out["kind_%d" % i] = np.frombuffer(kind.encode(), np.uint8)
        print(i, S.LAYOUT_NAMES[layout], w, h, slices, context, kind, slicecrc, len(rec), len(pkt))
    out["meta"] = np.array(meta, np.int64)
    np.savez_compressed(os.path.join(HERE, "ffv1_golden.npz"), **out)


if __name__ == "__main__":
    main()
