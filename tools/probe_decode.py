"""Decode + compare throughput probe (GPU box): encodes B frames on the device, then runs the `--check` kernel on the
encoder's arena against the device input. usage: probe_decode.py B kind [spw ...]; geometry from PROBE_W/H/LAYOUT/SLICES."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rawcooked_b200 import ffv1, ffv1dec, synth as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kind = sys.argv[2] if len(sys.argv) > 2 else "grain"
spws = [int(v) for v in sys.argv[3:]] or [4]
w, h, layout = int(os.environ.get("PROBE_W", 3840)), int(os.environ.get("PROBE_H", 2160)), int(os.environ.get("PROBE_LAYOUT", S.DPX_RGB_16_BE))
slices = int(os.environ.get("PROBE_SLICES", 24))
uniq = [S.synth_payload(w, h, layout, 3000 + k, kind) for k in range(2)]
d = torch.stack([torch.from_numpy(uniq[k % 2]) for k in range(B)]).cuda()
enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=B)
enc.encode_device(d.data_ptr(), B, 0)
torch.cuda.synchronize()
arena, off, ln = enc.packets_device(B)
for spw in spws:
    dec = ffv1dec.FFV1Decoder(w, h, layout, enc.config_record, max_frames=B, slices_per_warp=spw)
    for it in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        dec.decode_device(arena, off, ln, d_sources=d.data_ptr())
        mm, st = dec.result(B)
        dt = time.perf_counter() - t
    s = dec.stats()
    print("decode B=%d %s spw=%d: %.1f ms  %.1f fps  %.1f MPix/s  kernel %.1f ms  mismatch %d status %d  (%.0f cycles/sample at 1.965 GHz)" % (
        B, kind, spw, dt * 1e3, B / dt, B * w * h / dt / 1e6, s["decode_us"] / 1e3, sum(mm), max(st),
        s["decode_us"] * 1e-6 * 1.965e9 / (s["samples"] / (B * slices))))
    dec.close()
