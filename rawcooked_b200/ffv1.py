"""Host-side mirror of the encoder boundary: the option names are the ffmpeg options RAWcooked emits for the
FFV1 track (`-slices`, `-context`, `-coder`, `-slicecrc`, `-level 3`, `-g 1`;
/root/reference/Source/CLI/Global.cpp:938-989, Source/CLI/Output.cpp:41-57), the work is done by
libb200enc.so (include/b200enc.h) on a B200. No CPU fallback: a missing library or device raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb200enc.so")
_lib = None


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200enc error %d: %s" % (code, msg))
        self.code = code


class _Cfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("layout", C.c_int32), ("slices", C.c_int32),
                ("context", C.c_int32), ("coder", C.c_int32), ("slicecrc", C.c_int32), ("max_frames", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32 * 7)]


def load_library():
    """Loads the in-tree C-ABI library; raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError("rawcooked_b200/libb200enc.so is missing: run `python __graft_entry__.py` (there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    L.b200_last_error.restype = C.c_char_p
    L.b200_version.restype = C.c_uint32
    L.b200_ffv1_slice_grid.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.b200_ffv1_frame_bytes.restype = C.c_size_t
    L.b200_ffv1_frame_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_int32]
    L.b200_ffv1_open.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_void_p)]
    L.b200_ffv1_close.argtypes = [C.c_void_p]
    L.b200_ffv1_close.restype = None
    L.b200_ffv1_config_record.restype = C.c_size_t
    L.b200_ffv1_config_record.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.b200_ffv1_config_record_for.restype = C.c_size_t
    L.b200_ffv1_config_record_for.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_size_t]
    L.b200_ffv1_max_packet_bytes.restype = C.c_size_t
    L.b200_ffv1_max_packet_bytes.argtypes = [C.c_void_p]
    L.b200_ffv1_encode_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.b200_ffv1_submit_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]
    L.b200_ffv1_prefetch_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]
    L.b200_ffv1_encode_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.b200_ffv1_packets_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int32]
    L.b200_ffv1_fetch_packets.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int32]
    L.b200_ffv1_packet_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(C.c_size_t)]
    L.b200_ffv1_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.b200_ffv1_set_timing.argtypes = [C.c_void_p, C.c_int32]
    L.b200_ffv1_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200Error(rc, load_library().b200_last_error().decode())


def slice_grid(width, height, slices):
    L = load_library()
    nh, nv = C.c_int32(0), C.c_int32(0)
    _check(L.b200_ffv1_slice_grid(width, height, slices, C.byref(nh), C.byref(nv)))
    return nh.value, nv.value


def config_record(width, height, layout, slices=0, context=1, slicecrc=1):
    """FFV1 ConfigurationRecord (Matroska CodecPrivate) for this option set, computed on the host: no device needed."""
    L = load_library()
    cfg = _Cfg(width, height, layout, slices, context, 1, slicecrc, 1, 0)
    n = L.b200_ffv1_config_record_for(C.byref(cfg), None, 0)
    if n == 0:
        raise B200Error(-1, L.b200_last_error().decode())
    buf = C.create_string_buffer(n)
    L.b200_ffv1_config_record_for(C.byref(cfg), buf, n)
    return buf.raw[:n]


def frame_bytes(width, height, layout):
    return load_library().b200_ffv1_frame_bytes(width, height, layout)


class FFV1Encoder:
    """One FFV1 v3 video track encoder on one GPU; `encode` takes a batch of frame payloads (bytes / uint8 arrays in
    the file's own layout) and returns one packet (Matroska SimpleBlock payload) per frame."""

    def __init__(self, width, height, layout, slices=0, context=1, coder=1, slicecrc=1, level=3, g=1, max_frames=8, device=0):
        if level != 3 or g != 1:
            raise ValueError("RAWcooked's option set is -level 3 -g 1")
        self._L = load_library()
        cfg = _Cfg(width, height, layout, slices, context, coder, slicecrc, max_frames, device)
        h = C.c_void_p()
        _check(self._L.b200_ffv1_open(C.byref(cfg), C.byref(h)))
        self._h = h
        self.width, self.height, self.layout, self.max_frames, self.device = width, height, layout, max_frames, device
        self.frame_bytes = frame_bytes(width, height, layout)
        self.grid = slice_grid(width, height, slices)
        info = (C.c_int32 * 8)()
        _check(self._L.b200_ffv1_info(self._h, info))
        self.nbands, self.band_rows, self.nslices = info[2], info[3], info[4]
        n = self._L.b200_ffv1_config_record(self._h, None, 0)
        buf = C.create_string_buffer(n)
        self._L.b200_ffv1_config_record(self._h, buf, n)
        self.config_record = buf.raw[:n]
        self.max_packet_bytes = self._L.b200_ffv1_max_packet_bytes(self._h)

    def close(self):
        if self._h:
            self._L.b200_ffv1_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, frames):
        """frames: list of bytes-like / uint8 numpy arrays, each frame_bytes long (host memory). Returns list of bytes."""
        n = len(frames)
        arrs = [np.ascontiguousarray(np.frombuffer(f, np.uint8) if not isinstance(f, np.ndarray) else f, dtype=np.uint8) for f in frames]
        for a in arrs:
            if a.size != self.frame_bytes:
                raise ValueError("frame has %d bytes, layout needs %d" % (a.size, self.frame_bytes))
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        cap = min(self.max_packet_bytes, self.frame_bytes * 3 + (1 << 16)) * n
        out = np.empty(cap, np.uint8)
        off = (C.c_size_t * n)()
        ln = (C.c_size_t * n)()
        _check(self._L.b200_ffv1_encode_host(self._h, ptrs, n, out.ctypes.data, cap, off, ln))
        return [out[off[i]:off[i] + ln[i]].tobytes() for i in range(n)]

    def submit(self, frames):
        """Asynchronous host entry point: enqueues the batch (host->device copies band by band + kernels) and returns; up to
        two batches may be in flight, `fetch_packets` always collects the oldest. The arrays are kept alive until then."""
        n = len(frames)
        arrs = [np.ascontiguousarray(np.frombuffer(f, np.uint8) if not isinstance(f, np.ndarray) else f, dtype=np.uint8) for f in frames]
        for a in arrs:
            if a.size != self.frame_bytes:
                raise ValueError("frame has %d bytes, layout needs %d" % (a.size, self.frame_bytes))
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        self._inflight = getattr(self, "_inflight", [])[-1:] + [(arrs, ptrs)]
        _check(self._L.b200_ffv1_submit_host(self._h, ptrs, n))

    def prefetch(self, frames):
        """Starts the upload of the batch the NEXT submit will take (call it before fetching the previous batch's packets)."""
        n = len(frames)
        arrs = [np.ascontiguousarray(np.frombuffer(f, np.uint8) if not isinstance(f, np.ndarray) else f, dtype=np.uint8) for f in frames]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        self._prefetched = (arrs, ptrs)
        _check(self._L.b200_ffv1_prefetch_host(self._h, ptrs, n))

    def encode_device(self, d_ptr, n_frames, stream=0):
        """Asynchronous device-resident encode: d_ptr = device address of n_frames payloads back to back."""
        _check(self._L.b200_ffv1_encode_device(self._h, C.c_void_p(d_ptr), n_frames, C.c_void_p(stream)))

    def packets_device(self, n_frames):
        """After a stream sync: (arena device pointer, offsets, lengths)."""
        arena = C.c_void_p()
        off = (C.c_size_t * n_frames)()
        ln = (C.c_size_t * n_frames)()
        _check(self._L.b200_ffv1_packets_device(self._h, C.byref(arena), off, ln, n_frames))
        return arena.value, list(off), list(ln)

    def fetch_packets(self, n_frames, out=None):
        cap = min(self.max_packet_bytes, self.frame_bytes * 3 + (1 << 16)) * n_frames
        if out is None:
            out = np.empty(cap, np.uint8)
        off = (C.c_size_t * n_frames)()
        ln = (C.c_size_t * n_frames)()
        _check(self._L.b200_ffv1_fetch_packets(self._h, out.ctypes.data, out.size, off, ln, n_frames))
        return [out[off[i]:off[i] + ln[i]] for i in range(n_frames)]

    def stats(self):
        s = (C.c_uint64 * 8)()
        _check(self._L.b200_ffv1_stats(self._h, s))
        return {"launches": s[0], "bins": s[1], "samples": s[2], "packet_bytes": s[3],
                "model_us": s[4], "range_us": s[5], "pack_us": s[6], "emit_us": s[7]}

    def set_timing(self, enabled):
        _check(self._L.b200_ffv1_set_timing(self._h, 1 if enabled else 0))
