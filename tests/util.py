"""ctypes access to the checkers (oracle/ — test infrastructure) for the tests."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OracleCfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("layout", C.c_int32), ("num_h", C.c_int32),
                ("num_v", C.c_int32), ("context", C.c_int32), ("ec", C.c_int32)]


_oracle = None


def oracle():
    """The C restatement oracle/ffv1_oracle.c (built on demand)."""
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "libffv1_oracle.so")
        src = os.path.join(ORACLE_DIR, "ffv1_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libffv1_oracle.so"], stdout=subprocess.DEVNULL)
        O = C.CDLL(so)
        O.ffv1o_config_record.restype = C.c_size_t
        O.ffv1o_config_record.argtypes = [C.POINTER(OracleCfg), C.c_void_p, C.c_size_t]
        O.ffv1o_encode_frame.restype = C.c_size_t
        O.ffv1o_encode_frame.argtypes = [C.POINTER(OracleCfg), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        O.ffv1o_slice_bins.restype = C.c_size_t
        O.ffv1o_slice_bins.argtypes = [C.POINTER(OracleCfg), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        O.ffv1o_frame_bytes.restype = C.c_size_t
        O.ffv1o_frame_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        O.ffv1o_crc32.restype = C.c_uint32
        O.ffv1o_crc32.argtypes = [C.c_void_p, C.c_size_t]
        O.ffv1o_slice_grid.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        O.ffv1o_decode_frame.restype = C.c_int
        O.ffv1o_decode_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_uint32)]
        _oracle = O
    return _oracle


def oracle_grid(width, height, slices, bits):
    nh, nv = C.c_int(0), C.c_int(0)
    r = oracle().ffv1o_slice_grid(width, height, slices, bits, C.byref(nh), C.byref(nv))
    if r:
        raise ValueError("no slice grid for %d" % slices)
    return nh.value, nv.value


def oracle_record(width, height, layout, num_h, num_v, context=1, ec=1):
    cfg = OracleCfg(width, height, layout, num_h, num_v, context, ec)
    buf = C.create_string_buffer(4096)
    n = oracle().ffv1o_config_record(C.byref(cfg), buf, 4096)
    return buf.raw[:n]


def oracle_encode(payload, width, height, layout, num_h, num_v, context=1, ec=1, want_sizes=False):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    cfg = OracleCfg(width, height, layout, num_h, num_v, context, ec)
    cap = payload.size * 3 + 65536
    out = np.empty(cap, np.uint8)
    sizes = np.zeros(num_h * num_v, np.uint32)
    bins = C.c_uint64(0)
    n = oracle().ffv1o_encode_frame(C.byref(cfg), payload.ctypes.data, out.ctypes.data, cap, sizes.ctypes.data, C.byref(bins))
    assert n, "oracle overflow"
    pkt = out[:n].tobytes()
    if want_sizes:
        return pkt, sizes, bins.value
    return pkt


def oracle_decode(record, packet, width, height, layout):
    """The C restatement of the reference decoder (oracle/ffv1_oracle.c, second half): -> (payload bytes, status flags)."""
    n = oracle().ffv1o_frame_bytes(width, height, layout)
    out = np.zeros(n, np.uint8)
    flags = C.c_uint32(0)
    rc = oracle().ffv1o_decode_frame(bytes(record), len(record), width, height, layout, bytes(packet), len(packet), out.ctypes.data, n, C.byref(flags))
    if rc:
        raise RuntimeError("oracle decoder refused the stream")
    return out.tobytes(), flags.value


def oracle_slice_bins(payload, width, height, layout, num_h, num_v, sx, sy, context=1):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    cfg = OracleCfg(width, height, layout, num_h, num_v, context, 1)
    cap = (width // num_h + 1) * (height // num_v + 1) * 3 * 36 + 1024
    out = np.empty(cap, np.uint16)
    n = oracle().ffv1o_slice_bins(C.byref(cfg), payload.ctypes.data, sx, sy, out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


_ref = None


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_ffv1dec.so"))


def ref_decoder():
    """The UNMODIFIED reference decoder (oracle/_ref, built by oracle/build_ref.sh)."""
    global _ref
    if _ref is None:
        R = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_ffv1dec.so"))
        R.ref_ffv1_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t, C.c_int]
        R.ref_crc32.restype = C.c_uint32
        R.ref_crc32.argtypes = [C.c_void_p, C.c_size_t]
        _ref = R
    return _ref


def ref_decode(record, packet, width, height, layout, threads=1):
    """Decode one packet with the reference decoder into the file payload layout. Returns bytes; raises on decoder error.
    The reference rounds the plane buffer up to 32 bits (RawFrame.h:62-68); that (uninitialised) tail is trimmed."""
    container, flavor = (0, layout) if layout < 32 else (1, {32: 0, 33: 1, 34: 2}[layout])
    cap = width * height * 8 + 4096
    out = np.zeros(cap, np.uint8)
    osz = C.c_size_t(0)
    err = C.create_string_buffer(256)
    rc = ref_decoder().ref_ffv1_decode(record, len(record), width, height, container, flavor, packet, len(packet),
                                       out.ctypes.data, cap, C.byref(osz), err, 256, threads)
    if rc:
        raise RuntimeError("reference decoder: rc=%d %s" % (rc, err.value.decode()))
    from rawcooked_b200 import synth as S
    n = S.frame_bytes(width, height, layout)
    # 8-bit and 16-bit DPX files pad every line to 32 bits (DPX.cpp:478-482, what the reference's parser sizes the file by and
    # what ffmpeg's dpx decoder skips), but the decoder's plane holds whole 3- or 6-byte blocks back to back (RawFrame.h:118-131:
    # a line is only aligned when it ends in an incomplete block). The padding is not pixel data: re-insert it as zeros so that
    # the result compares with a zero-padded source payload.
    valid = {S.DPX_RGB_8: 3 * width, S.DPX_RGB_16_LE: 6 * width, S.DPX_RGB_16_BE: 6 * width}.get(layout)
    rb = S.row_bytes(width, layout)
    if valid is not None and valid != rb:
        assert osz.value - valid * height in range(0, 4), (osz.value, valid * height)
        rows = np.zeros((height, rb), np.uint8)
        rows[:, :valid] = out[:valid * height].reshape(height, valid)
        return rows.tobytes()
    assert osz.value - n in range(0, 4), (osz.value, n)
    return out[:n].tobytes()


def ref_rawcooked():
    p = os.path.join(ORACLE_DIR, "_ref", "rawcooked")
    return p if os.path.exists(p) else None


_golden = None


def golden():
    """tests/golden/ffv1_golden.npz: libavcodec's own records and packets (tests/golden/make_golden.py)."""
    global _golden
    if _golden is None:
        _golden = np.load(os.path.join(ROOT, "tests", "golden", "ffv1_golden.npz"))
    return _golden


def golden_count():
    return len(golden()["meta"])


def golden_case(i):
    """(w, h, layout, slices, context, slicecrc, payload, record, packet) of golden case i."""
    from rawcooked_b200 import synth as S
    G = golden()
    w, h, layout, slices, context, seed, slicecrc = (int(v) for v in G["meta"][i])
    kind = G["kind_%d" % i].tobytes().decode()
    payload = G["payload_%d" % i]
    if payload.size == 0:
        payload = S.synth_payload(w, h, layout, seed, kind)
    return w, h, layout, slices, context, slicecrc, payload, G["record_%d" % i].tobytes(), G["packet_%d" % i].tobytes()
