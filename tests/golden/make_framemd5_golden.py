"""Generates tests/golden/framemd5_golden.json: the framemd5 files FFmpeg's own libraries produce for seeded image
sequences — every frame decoded by libavcodec's dpx / tiff decoder (oracle/avcodec_rawframe.py), hashed and written by
libavformat's framemd5 muxer (oracle/avformat_framemd5.py). This is the second output `rawcooked --framemd5` asks the
encoder process for (/root/reference/Source/CLI/Output.cpp:312-332). Run in the build container (needs the bundled FFmpeg
libraries); the fixture travels, the libraries need not."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import avcodec_rawframe as R  # noqa: E402
import avformat_framemd5 as M  # noqa: E402
from rawcooked_b200 import synth as S  # noqa: E402

CASES = [  # w, h, layout, first seed, frames, fps_num, fps_den
    (64, 48, 0, 100, 3, 24, 1), (37, 11, 0, 110, 2, 25, 1), (96, 64, 1, 120, 2, 24, 1), (96, 64, 2, 130, 3, 30000, 1001),
    (40, 20, 3, 140, 2, 24, 1), (264, 100, 4, 150, 2, 24, 1), (37, 11, 4, 155, 2, 24, 1), (40, 20, 5, 160, 2, 24, 1),
    (64, 48, 6, 170, 2, 24, 1), (131, 17, 7, 180, 3, 24, 1), (64, 48, 32, 190, 2, 24, 1), (64, 48, 33, 200, 2, 24, 1),
    (33, 9, 34, 210, 2, 24, 1),
]
out = []
for w, h, layout, seed, n, fn, fd in CASES:
    raws = []
    fmt = None
    for i in range(n):
        payload = S.synth_payload(w, h, layout, seed + i)
        f = S.dpx_file(w, h, layout, payload, i) if layout < 32 else S.tiff_file(w, h, layout, payload)
        fmt, W, H, raw = R.decode_image(f, "dpx" if layout < 32 else "tiff")
        assert (W, H) == (w, h)
        raws.append(raw)
    out.append({"w": w, "h": h, "layout": layout, "seed": seed, "frames": n, "fps_num": fn, "fps_den": fd, "pix_fmt": fmt,
                "text": M.framemd5(raws, w, h, fmt, fn, fd)})
json.dump({"libavformat": "62.3.100", "cases": out}, open(os.path.join(ROOT, "tests", "golden", "framemd5_golden.json"), "w"), indent=1)
print("wrote", len(out), "cases")
