import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rawcooked_b200 import ffv1, synth as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kind = sys.argv[2] if len(sys.argv) > 2 else "grain"
w, h, layout = int(os.environ.get("PROBE_W", 3840)), int(os.environ.get("PROBE_H", 2160)), int(os.environ.get("PROBE_LAYOUT", S.DPX_RGB_16_BE))
slices = int(os.environ.get("PROBE_SLICES", 24))
uniq = [S.synth_payload(w, h, layout, 3000 + k, kind) for k in range(2)]
d = torch.stack([torch.from_numpy(uniq[k % 2]) for k in range(B)]).cuda()
enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=B)
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    enc.encode_device(d.data_ptr(), B, 0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("B=%d %s: %.1f ms  %.1f fps  %.1f MPix/s" % (B, kind, dt * 1e3, B / dt, B * w * h / dt / 1e6))
arena, off, ln = enc.packets_device(B)
if os.environ.get("PROBE_KERNELS"):
    enc.set_timing(True)
    enc.encode_device(d.data_ptr(), B, 0)
    torch.cuda.synchronize()
    enc.packets_device(B)
    ts = enc.stats()
    print("kernel ms per batch: model %.1f range %.1f emit %.1f pack %.1f" % (ts["model_us"] / 1e3, ts["range_us"] / 1e3, ts["emit_us"] / 1e3, ts["pack_us"] / 1e3))
    enc.set_timing(False)
st = enc.stats()
print(st, "bins/sample", st["bins"] / st["samples"], "ratio", st["packet_bytes"] / (B * enc.frame_bytes))
