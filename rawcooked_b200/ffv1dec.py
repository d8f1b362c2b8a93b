"""Host-side mirror of the decode + compare boundary (include/b200dec.h): what `rawcooked --check` does per video packet
through ffv1_wrapper (/root/reference/Source/Lib/CoDec/Wrapper.cpp:71-128) and FileWriter's compare
(Source/Lib/Utils/FileIO/FileWriter.cpp:464-727), done by libb200enc.so on a B200. No CPU fallback."""
import ctypes as C

import numpy as np

from . import ffv1 as _enc

BAD_TAIL, BAD_CRC, BAD_HEADER, UNDERRUN, JUNK, ERROR_STATUS = 1, 2, 4, 8, 16, 32


class _DecCfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("layout", C.c_int32), ("max_frames", C.c_int32),
                ("device", C.c_int32), ("slices_per_warp", C.c_int32), ("reserved", C.c_int32 * 6)]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("version", "micro_version", "coder_type", "colorspace_type", "bits_per_raw_sample",
                                         "chroma_planes", "log2_h_chroma_subsample", "log2_v_chroma_subsample", "alpha_plane",
                                         "num_h_slices", "num_v_slices", "quant_table_set_count", "ec", "intra")] + \
               [("context_count", C.c_int32 * 8), ("crc_ok", C.c_int32)]


_bound = False


def _lib():
    global _bound
    L = _enc.load_library()
    if not _bound:
        L.b200_ffv1_parse_config_record.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Params)]
        L.b200_ffv1_dec_open.argtypes = [C.POINTER(_DecCfg), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.b200_ffv1_dec_close.argtypes = [C.c_void_p]
        L.b200_ffv1_dec_close.restype = None
        L.b200_ffv1_decode_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_uint32)]
        L.b200_ffv1_check_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.b200_ffv1_decode_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        L.b200_ffv1_dec_result.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_int32]
        L.b200_ffv1_dec_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        _bound = True
    return L


def parse_config_record(record):
    """parameters::Parse (FFV1_Parameters.cpp:23-183) on the host; raises B200Error with the reference's error text."""
    p = Params()
    _enc._check(_lib().b200_ffv1_parse_config_record(record, len(record), C.byref(p)))
    return p


class FFV1Decoder:
    """One V_FFV1 track: CodecPrivate -> decoder; packets -> payloads in the source file's byte layout, or a compare."""

    def __init__(self, width, height, layout, record, max_frames=8, device=0, slices_per_warp=0):
        self._L = _lib()
        self._h = C.c_void_p()
        cfg = _DecCfg(width, height, layout, max_frames, device, slices_per_warp)
        _enc._check(self._L.b200_ffv1_dec_open(C.byref(cfg), record, len(record), C.byref(self._h)))
        self.frame_bytes = self._L.b200_ffv1_frame_bytes(width, height, layout)
        self.max_frames = max_frames

    def close(self):
        if self._h:
            self._L.b200_ffv1_dec_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    @staticmethod
    def _ptrs(bufs):
        keep = [np.frombuffer(b, np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, np.uint8) for b in bufs]
        arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
        return keep, arr

    def decode(self, packets):
        """-> (list of payload bytes, list of status bits)"""
        n = len(packets)
        keep, pp = self._ptrs(packets)
        lens = (C.c_size_t * n)(*[len(p) for p in packets])
        outs = [np.empty(self.frame_bytes, np.uint8) for _ in range(n)]
        op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        st = (C.c_uint32 * n)()
        _enc._check(self._L.b200_ffv1_decode_host(self._h, pp, lens, n, op, st))
        return [o.tobytes() for o in outs], list(st)

    def check(self, packets, sources):
        """-> (mismatch counts, status bits): the --check operation, compare on the GPU"""
        n = len(packets)
        keep, pp = self._ptrs(packets)
        keep2, sp = self._ptrs(sources)
        for s in keep2:
            assert s.size == self.frame_bytes
        lens = (C.c_size_t * n)(*[len(p) for p in packets])
        mm = (C.c_uint64 * n)()
        st = (C.c_uint32 * n)()
        _enc._check(self._L.b200_ffv1_check_host(self._h, pp, lens, n, sp, mm, st))
        return list(mm), list(st)

    def decode_device(self, d_packets, offs, lens, d_out=None, d_sources=None, stream=None):
        n = len(offs)
        o = (C.c_size_t * n)(*offs)
        l = (C.c_size_t * n)(*lens)
        _enc._check(self._L.b200_ffv1_decode_device(self._h, d_packets, o, l, n, d_out, d_sources, stream))

    def result(self, n):
        mm = (C.c_uint64 * n)()
        st = (C.c_uint32 * n)()
        _enc._check(self._L.b200_ffv1_dec_result(self._h, mm, st, n))
        return list(mm), list(st)

    def stats(self):
        s = (C.c_uint64 * 8)()
        _enc._check(self._L.b200_ffv1_dec_stats(self._h, s))
        return {"index_us": s[0], "decode_us": s[1], "total_us": s[2], "slices": s[3], "samples": s[4]}
