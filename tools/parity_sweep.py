import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import util
from rawcooked_b200 import ffv1, synth as S
cases = []
for layout in (S.DPX_RGB_16_BE, S.DPX_RGB_10_FA_BE, S.DPX_RGB_8):
    for kind in ("const", "flat", "grain", "white"):
        for (w, h, sl) in ((64, 32, 4), (96, 64, 4), (200, 150, 6), (1300, 40, 4)):
            if layout == S.DPX_RGB_8 and (w * 3) % 4: continue
            cases.append((layout, kind, w, h, sl, 1))
cases.append((S.DPX_RGB_16_BE, "grain", 96, 64, 4, 0))
for kind in ("grain", "flat", "const", "white"):
    cases.append((S.DPX_RGB_8, kind, 640, 480, 16, 1))
    cases.append((S.DPX_RGB_10_FA_BE, kind, 2048, 160, 4, 1))
for layout, kind, w, h, sl, ctx in cases:
    try:
        enc = ffv1.FFV1Encoder(w, h, layout, slices=sl, context=ctx, max_frames=2)
    except Exception as e:
        print("OPENFAIL", layout, kind, w, h, sl, e); continue
    try:
        nh, nv = enc.grid
        f = S.synth_payload(w, h, layout, 40, kind)
        try:
            p = enc.encode([f])[0]
        except Exception as e:
            print("ENCFAIL", layout, kind, w, h, sl, e); continue
        o = util.oracle_encode(f, w, h, layout, nh, nv, ctx)
        if p == o:
            print("ok  ", layout, kind, w, h, sl, ctx, len(p))
        else:
            n = min(len(p), len(o))
            d = next((i for i in range(n) if p[i] != o[i]), n)
            print("DIFF", layout, kind, w, h, sl, ctx, "len", len(p), len(o), "first diff at", d)
    finally:
        enc.close()
