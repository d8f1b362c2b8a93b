"""Whole front-end throughput: N synthetic 4K 16-bit DPX files in /dev/shm -> b200enc (argv as RAWcooked emits it) -> MKV in
/dev/shm. Includes file reads (threaded, double-buffered), H2D, encode, D2H, Matroska write. usage: cli_throughput.py [N]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rawcooked_b200 import synth as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w, h, layout = 3840, 2160, S.DPX_RGB_16_BE
d = "/dev/shm/b200_cli_tp"
os.makedirs(d, exist_ok=True)
uniq = [S.synth_payload(w, h, layout, 5000 + k, "grain") for k in range(4)]
for i in range(n):
    p = os.path.join(d, "f_%06d.dpx" % i)
    if not os.path.exists(p):
        open(p, "wb").write(S.dpx_file(w, h, layout, uniq[i % 4], i))
# what the box's tmpfs itself can take: the front-end's steady state is bound by it (one writer, then 8 in parallel)
import threading
blob = np.random.default_rng(0).integers(0, 256, 1 << 30, dtype=np.uint8).tobytes()
t = time.perf_counter()
with open(os.path.join(d, "probe.bin"), "wb") as f:
    f.write(blob)
print("tmpfs write, 1 thread: %.2f GB/s" % (len(blob) / (time.perf_counter() - t) / 1e9))
def _w(i):
    with open(os.path.join(d, "probe%d.bin" % i), "wb") as f:
        f.write(blob[: 1 << 28])
ths = [threading.Thread(target=_w, args=(i,)) for i in range(8)]
t = time.perf_counter()
[x.start() for x in ths]; [x.join() for x in ths]
print("tmpfs write, 8 threads / 8 files: %.2f GB/s" % (8 * (1 << 28) / (time.perf_counter() - t) / 1e9))
for f in os.listdir(d):
    if f.startswith("probe"):
        os.remove(os.path.join(d, f))
del blob
out = os.path.join(d, "out.mkv")
cmd = [os.path.join(ROOT, "rawcooked_b200", "b200enc"), "-xerror", "-framerate", "24", "-r", "24", "-f", "image2", "-c:v", "dpx", "-start_number", "000000",
       "-i", os.path.join(d, "f_%06d.dpx"), "-c:a", "flac", "-c:v", "ffv1", "-coder", "1", "-context", "1", "-f", "matroska", "-g", "1", "-level", "3",
       "-slicecrc", "1", "-slices", "24", "-y", "-f", "matroska", out]
# plain, then with the encode-time verification (every packet decoded again on the GPU), then with a framemd5 second output
variants = [("plain", {}, []), ("plain", {}, [])]
if len(sys.argv) > 2:
    variants += [("B200_VERIFY=1", {"B200_VERIFY": "1"}, []), ("-f framemd5", {}, ["-f", "framemd5", os.path.join(d, "out.framemd5")])]
for it, (label, env, extra) in enumerate(variants):
    t = time.perf_counter()
    r = subprocess.run(cmd + extra, capture_output=True, text=True, env=dict(os.environ, B200_CLI_TIMING='1', **env))
    dt = time.perf_counter() - t
    print("run %d (%s): rc=%d %.2f s  %.1f fps  %.0f MPix/s  (process start + encoder open included)  mkv %.1f MB %s" %
          (it, label, r.returncode, dt, n / dt, n * w * h / dt / 1e6, os.path.getsize(out) / 1e6 if os.path.exists(out) else 0, '\n' + r.stderr[-700:]))
for f in os.listdir(d):
    os.remove(os.path.join(d, f))
