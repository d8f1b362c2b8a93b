"""Analysis-pass probe (GPU box): MD5 of B files and the padding-bit test of B payloads already in device memory, next to
the reference's own md5.c on one core (what `rawcooked --hash` does per file). usage: probe_scan.py [B]"""
import sys, time, os, ctypes as C, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from rawcooked_b200 import scan, synth as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w, h = 3840, 2160
for layout, name in ((S.DPX_RGB_16_BE, "16-bit"), (S.DPX_RGB_10_FA_BE, "10-bit Filled A")):
    fb = S.frame_bytes(w, h, layout)
    one = torch.from_numpy(np.array(S.synth_payload(w, h, layout, 1), np.uint8).reshape(-1))
    d = one.cuda().repeat(B)
    d[fb * (B // 2) + 1003] |= 3
    sc = scan.Scanner(max_items=B, max_bytes=0)
    offs = [i * fb for i in range(B)]
    for it in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        dig = sc.md5_device(d.data_ptr(), offs, [fb] * B)
        dt = time.perf_counter() - t
    want = hashlib.md5(one.numpy().tobytes()).digest()
    ok = all(g == want for i, g in enumerate(dig) if i != B // 2) and dig[B // 2] != want
    print("md5 %s: %d x %.1f MB in %.1f ms = %.1f GB/s = %.0f files/s (kernel %.1f ms) digests ok: %s" % (
        name, B, fb / 1e6, dt * 1e3, B * fb / dt / 1e9, B / dt, sc.stats()["kernel_us"] / 1e3, ok))
    for it in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        cnt, first = sc.padding_device(w, h, layout, d.data_ptr(), B)
        dt = time.perf_counter() - t
    print("padding %s: %.2f ms = %.0f GB/s  nonzero %d first %s (kernel %.3f ms, %.1f MB read)" % (
        name, dt * 1e3, sc.stats()["bytes"] / dt / 1e9, sum(cnt), [f for f in first if f is not None], sc.stats()["kernel_us"] / 1e3, sc.stats()["bytes"] / 1e6))
    sc.close()
    del d
import util
if util.ref_available():
    R = util.ref_decoder()
    R.ref_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    buf = one.numpy()
    out = C.create_string_buffer(16)
    t = time.perf_counter()
    for _ in range(3):
        R.ref_md5(buf.ctypes.data, buf.size, out)
    dt = (time.perf_counter() - t) / 3
    print("reference md5.c, one core: %.1f MB in %.1f ms = %.2f GB/s = %.1f files/s" % (buf.size / 1e6, dt * 1e3, buf.size / dt / 1e9, 1 / dt))
