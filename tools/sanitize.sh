#!/bin/bash
# compute-sanitizer over a small encode of every kernel variant (16-bit replicated table, 10-bit, 8-bit compact, FLAC):
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
    enc.close()
print("sanitize case ok")
PY
for TOOL in memcheck racecheck initcheck; do
  echo "== $TOOL"
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 5 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|Error|Invalid|Uninit" | head -12
done
