"""Host-side mirror of the analysis-pass helpers (include/b200scan.h): what RAWcooked's input parsers compute per source
file before the encoder runs — `--hash` (input_base::Hash, /root/reference/Source/Lib/Utils/FileIO/Input_Base.cpp:54-81) and
`--check-padding` (dpx::ParseBuffer, Source/Lib/Uncompressed/DPX/DPX.cpp:500-608) — done by libb200enc.so on a B200."""
import ctypes as C

import numpy as np

from . import ffv1 as _enc

_bound = False
NONE = 0xFFFFFFFFFFFFFFFF


def _lib():
    global _bound
    L = _enc.load_library()
    if not _bound:
        L.b200_scan_open.argtypes = [C.c_int32, C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]
        L.b200_scan_close.argtypes = [C.c_void_p]
        L.b200_scan_close.restype = None
        L.b200_md5_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int32, C.c_void_p]
        L.b200_md5_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int32, C.c_void_p, C.c_void_p]
        L.b200_padding_host.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.b200_padding_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_void_p, C.c_int32,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        L.b200_scan_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.b200_rawvideo_bytes.restype = C.c_size_t
        L.b200_rawvideo_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_int32]
        L.b200_rawvideo_pix_fmt.restype = C.c_char_p
        L.b200_rawvideo_pix_fmt.argtypes = [C.c_int32]
        L.b200_framemd5_host.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]
        L.b200_framemd5_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _bound = True
    return L


def rawvideo_bytes(width, height, layout):
    return _lib().b200_rawvideo_bytes(width, height, layout)


def rawvideo_pix_fmt(layout):
    return _lib().b200_rawvideo_pix_fmt(layout).decode()


class Scanner:
    def __init__(self, max_items=128, max_bytes=1 << 28, device=0):
        self._L = _lib()
        self._h = C.c_void_p()
        _enc._check(self._L.b200_scan_open(device, max_items, max_bytes, C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.b200_scan_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    def md5(self, buffers):
        """MD5 digests (16 bytes each) of bytes-like objects lying in host memory."""
        n = len(buffers)
        keep = [np.frombuffer(b, np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, np.uint8) for b in buffers]
        ptrs = (C.c_void_p * n)(*[k.ctypes.data if k.size else 0 for k in keep])
        lens = (C.c_size_t * n)(*[k.size for k in keep])
        out = C.create_string_buffer(16 * n)
        _enc._check(self._L.b200_md5_host(self._h, ptrs, lens, n, out))
        return [out.raw[16 * i:16 * i + 16] for i in range(n)]

    def md5_device(self, d_base, offs, lens, stream=None):
        n = len(offs)
        o = (C.c_size_t * n)(*offs)
        l = (C.c_size_t * n)(*lens)
        out = C.create_string_buffer(16 * n)
        _enc._check(self._L.b200_md5_device(self._h, d_base, o, l, n, out, stream))
        return [out.raw[16 * i:16 * i + 16] for i in range(n)]

    def padding(self, width, height, layout, payloads, want_masked=False):
        """-> (nonzero counts, first offsets (None if all padding bits are zero), masked payloads or None)"""
        n = len(payloads)
        keep = [np.frombuffer(b, np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, np.uint8) for b in payloads]
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        cnt = (C.c_uint64 * n)()
        first = (C.c_uint64 * n)()
        outs = [np.empty(k.size, np.uint8) for k in keep] if want_masked else None
        mp = (C.c_void_p * n)(*[o.ctypes.data for o in outs]) if want_masked else None
        _enc._check(self._L.b200_padding_host(self._h, width, height, layout, ptrs, n, cnt, first, mp))
        return list(cnt), [None if f == NONE else int(f) for f in first], outs

    def padding_device(self, width, height, layout, d_payloads, n, d_masked=None, stream=None):
        cnt = (C.c_uint64 * n)()
        first = (C.c_uint64 * n)()
        _enc._check(self._L.b200_padding_device(self._h, width, height, layout, d_payloads, n, cnt, first, d_masked, stream))
        return list(cnt), [None if f == NONE else int(f) for f in first]

    def framemd5(self, width, height, layout, payloads):
        """MD5 of every frame as a `-f framemd5` output hashes it (FFmpeg's rawvideo form of the flavor)."""
        n = len(payloads)
        keep = [np.frombuffer(b, np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, np.uint8) for b in payloads]
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        out = C.create_string_buffer(16 * n)
        _enc._check(self._L.b200_framemd5_host(self._h, width, height, layout, ptrs, n, out))
        return [out.raw[16 * i:16 * i + 16] for i in range(n)]

    def stats(self):
        s = (C.c_uint64 * 4)()
        _enc._check(self._L.b200_scan_stats(self._h, s))
        return {"kernel_us": s[0], "bytes": s[1]}
