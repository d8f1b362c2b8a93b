"""Analysis-pass helpers (include/b200scan.h): MD5 of whole files as `--hash` computes it and the DPX
padding-bit test of `--check-padding`. CPU tests pin the oracle (oracle/scan_oracle.py against hand-made cases, the
reference's md5.c against hashlib); the GPU tests compare the CUDA path through the C ABI with both."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

import util
from rawcooked_b200 import synth as S

sys.path.insert(0, os.path.join(util.ROOT, "oracle"))
import scan_oracle  # noqa: E402


def ref_md5(data):
    R = util.ref_decoder()
    R.ref_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    out = C.create_string_buffer(16)
    buf = np.frombuffer(bytes(data), np.uint8)
    R.ref_md5(buf.ctypes.data if buf.size else None, buf.size, out)
    return out.raw


LENGTHS = [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1000, 4096, 65536 + 7, (1 << 20) + 13]


def test_reference_md5_is_rfc1321():
    if not util.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for n in LENGTHS[:14]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ref_md5(d) == hashlib.md5(d).digest()
    assert ref_md5(b"abc").hex() == "900150983cd24fb0d6963f7d28e17f72"       # RFC 1321 test suite


def _payload_with_padding(w, h, layout, seed, dirty):
    p = np.array(S.synth_payload(w, h, layout, seed), np.uint8).reshape(-1).copy()
    if dirty:
        rng = np.random.default_rng(seed)
        pos = rng.integers(0, p.size, 40)
        p[pos] |= rng.integers(1, 256, 40).astype(np.uint8)          # sets padding bits here and there (and sample bits: harmless)
    return p


def test_padding_oracle_on_hand_made_cases():
    # 10-bit Filled A big endian: the two low bits of every 32-bit word = bits 0..1 of the fourth byte (DPX.cpp:523-534)
    w, h = 4, 2
    p = np.zeros(4 * w * h, np.uint8)
    assert scan_oracle.padding_test(p, w, h, 2)[:2] == (0, None)
    p[7] = 0x02
    p[8] = 0xFF                                                        # sample bits only: not padding
    n, first, masked = scan_oracle.padding_test(p, w, h, 2)
    assert (n, first) == (1, 7) and masked[7] == 2 and masked.sum() == 2
    # little endian: the first byte of the word
    n, first, masked = scan_oracle.padding_test(p, w, h, 1)
    assert (n, first) == (1, 8) and masked[8] == 3
    # 12-bit Filled A: low nibble of every 16-bit sample
    q = np.zeros(6 * w * h, np.uint8)
    q[5] = 0x1F
    assert scan_oracle.padding_test(q, w, h, 5)[:2] == (1, 5)          # big endian: second byte
    assert scan_oracle.padding_test(q, w, h, 3)[:2] == (0, None)       # little endian: first byte
    # 12-bit packed, width 3: 108 bits per row -> 12 used bits in the last word, mask 0xFFFFF000 on the big-endian word
    r = np.zeros(scan_oracle.row_bytes(3, 4) * 2, np.uint8)
    assert scan_oracle.row_bytes(3, 4) == 16
    r[16 + 12] = 0x80
    n, first, masked = scan_oracle.padding_test(r, 3, 2, 4)
    assert (n, first) == (1, 28) and masked[28] == 0x80
    r[15] = 0xFF                                                       # low byte of row 0's last word: used bits 0..7 -> not padding
    assert scan_oracle.padding_test(r, 3, 2, 4)[0] == 1
    # widths whose rows end on a word boundary have nothing to test
    assert scan_oracle.padding_test(np.full(scan_oracle.row_bytes(8, 4) * 2, 255, np.uint8), 8, 2, 4)[:2] == (0, None)


def test_scan_needs_a_device():
    import torch
    from rawcooked_b200 import ffv1, scan
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ffv1.B200Error) as e:
        scan.Scanner()
    assert e.value.code == -2


@pytest.mark.gpu
def test_cuda_md5_equals_reference_md5():
    from rawcooked_b200 import scan
    rng = np.random.default_rng(11)
    bufs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in LENGTHS]
    bufs += [S.dpx_file(96, 64, S.DPX_RGB_10_FA_BE, S.synth_payload(96, 64, S.DPX_RGB_10_FA_BE, 9), 0) for _ in range(3)]   # whole files, as --hash
    sc = scan.Scanner(max_items=64, max_bytes=8 << 20)
    try:
        got = sc.md5(bufs)
        for b, g in zip(bufs, got):
            assert g == hashlib.md5(b).digest()
            if util.ref_available():
                assert g == ref_md5(b)
        # more messages than one warp, ragged lengths, and ranges that do not start on a 16-byte boundary (device entry point)
        torch = pytest.importorskip("torch")
        blob = rng.integers(0, 256, 300000, dtype=np.uint8)
        d = torch.from_numpy(blob).cuda()
        offs = [int(v) for v in rng.integers(0, 200000, 70)]
        lens = [int(v) for v in rng.integers(0, 90000, 70)]
        sc2 = scan.Scanner(max_items=70, max_bytes=0)
        got = sc2.md5_device(d.data_ptr(), offs, lens)
        for o, l, g in zip(offs, lens, got):
            assert g == hashlib.md5(blob[o:o + l].tobytes()).digest()
        sc2.close()
    finally:
        sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("w,h", [(96, 64), (131, 77), (3, 5)])
def test_cuda_padding_test_equals_oracle(layout, w, h):
    from rawcooked_b200 import scan
    sc = scan.Scanner(max_items=4, max_bytes=4 * S.frame_bytes(w, h, layout))
    try:
        payloads = [_payload_with_padding(w, h, layout, 60 + k, dirty=(k % 2 == 1)) for k in range(4)]
        cnt, first, masked = sc.padding(w, h, layout, payloads, want_masked=True)
        for k, p in enumerate(payloads):
            n, f, m = scan_oracle.padding_test(p, w, h, layout)
            assert (cnt[k], first[k]) == (n, f), (layout, w, h, k)
            assert np.array_equal(masked[k], m)
        cnt2, first2, _ = sc.padding(w, h, layout, payloads)
        assert (cnt2, first2) == (cnt, first)
    finally:
        sc.close()


# ---------------------------------------------------------------------------------------------------------------------
# framemd5 (the second output of RAWcooked's --framemd5, Output.cpp:312-332)
def _framemd5_golden():
    import json
    return json.load(open(os.path.join(util.ROOT, "tests", "golden", "framemd5_golden.json")))["cases"]


def _without_software(text):
    return "\n".join(l for l in text.splitlines() if not l.startswith("#software:")) + "\n"


def test_framemd5_oracle_equals_ffmpeg_golden():
    # the numpy restatement (pix_fmt per flavor, plane order, text layout) against the files FFmpeg's own libavcodec decoders
    # + libavformat framemd5 muxer produced (tests/golden/make_framemd5_golden.py)
    for c in _framemd5_golden():
        w, h, layout = c["w"], c["h"], c["layout"]
        assert scan_oracle.PIX_FMT[layout] == c["pix_fmt"]
        rows = []
        for i in range(c["frames"]):
            raw = scan_oracle.rawvideo_frame(S.synth_payload(w, h, layout, c["seed"] + i), w, h, layout)
            rows.append((len(raw), hashlib.md5(raw).digest()))
        assert scan_oracle.framemd5_text(w, h, c["fps_num"], c["fps_den"], rows) == _without_software(c["text"]), (w, h, layout)


@pytest.mark.gpu
def test_cuda_framemd5_equals_ffmpeg_golden_and_oracle():
    from rawcooked_b200 import scan
    for c in _framemd5_golden():
        w, h, layout, n = c["w"], c["h"], c["layout"], c["frames"]
        payloads = [S.synth_payload(w, h, layout, c["seed"] + i) for i in range(n)]
        sc = scan.Scanner(max_items=n, max_bytes=n * S.frame_bytes(w, h, layout))
        try:
            digs = sc.framemd5(w, h, layout, payloads)
            assert scan.rawvideo_pix_fmt(layout) == c["pix_fmt"]
            size = scan.rawvideo_bytes(w, h, layout)
            assert scan_oracle.framemd5_text(w, h, c["fps_num"], c["fps_den"], [(size, d) for d in digs]) == _without_software(c["text"])
        finally:
            sc.close()
    # a larger frame of every layout against the oracle
    for layout in sorted(S.LAYOUT_BITS):
        w, h = 333, 41
        payloads = [S.synth_payload(w, h, layout, 700 + i) for i in range(3)]
        sc = scan.Scanner(max_items=3, max_bytes=3 * S.frame_bytes(w, h, layout))
        try:
            assert sc.framemd5(w, h, layout, payloads) == [hashlib.md5(scan_oracle.rawvideo_frame(p, w, h, layout)).digest() for p in payloads]
        finally:
            sc.close()
