"""Host-side mirror of the FLAC encoder boundary (`-c:a flac`, /root/reference/Source/CLI/Global.cpp:949-950): the work is
done by libb200enc.so on a B200 (include/b200enc.h, b200_flac_*). No CPU fallback."""
import ctypes as C

import numpy as np

from .ffv1 import B200Error, load_library


class _Cfg(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits", C.c_uint32), ("block_size", C.c_int32),
                ("max_blocks", C.c_int32), ("device", C.c_int32), ("fixed_only", C.c_int32), ("reserved", C.c_int32 * 5)]


def _lib():
    L = load_library()
    if not hasattr(L, "_flac_ready"):
        L.b200_flac_open.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_void_p)]
        L.b200_flac_close.argtypes = [C.c_void_p]
        L.b200_flac_close.restype = None
        L.b200_flac_block_size.argtypes = [C.c_void_p]
        L.b200_flac_max_frame_bytes.restype = C.c_size_t
        L.b200_flac_max_frame_bytes.argtypes = [C.c_void_p]
        L.b200_flac_codec_private.restype = C.c_size_t
        L.b200_flac_codec_private.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]
        L.b200_flac_encode_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
        L._flac_ready = True
    return L


def pcm_to_wav_bytes(pcm, bits):
    """int32 [n, channels] -> the bytes of a WAV data chunk (8-bit unsigned, 16/24-bit signed little endian)."""
    n, ch = pcm.shape
    if bits == 8:
        return (pcm + 128).astype(np.uint8).reshape(-1)
    b = np.ascontiguousarray(pcm.astype("<i4")).view(np.uint8).reshape(n, ch, 4)[:, :, : bits // 8]
    return np.ascontiguousarray(b).reshape(-1)


class FLACEncoder:
    def __init__(self, sample_rate, channels, bits, block_size=0, max_blocks=256, device=0, fixed_only=False):
        self._L = _lib()
        cfg = _Cfg(sample_rate, channels, bits, block_size, max_blocks, device, 1 if fixed_only else 0)
        h = C.c_void_p()
        rc = self._L.b200_flac_open(C.byref(cfg), C.byref(h))
        if rc:
            raise B200Error(rc, self._L.b200_last_error().decode())
        self._h = h
        self.sample_rate, self.channels, self.bits, self.max_blocks = sample_rate, channels, bits, max_blocks
        self.block_size = self._L.b200_flac_block_size(h)
        self.max_frame_bytes = self._L.b200_flac_max_frame_bytes(h)
        self._frames = 0
        self._samples = 0

    def close(self):
        if self._h:
            self._L.b200_flac_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, wav_bytes):
        """wav_bytes: uint8 array / bytes of interleaved PCM as in the WAV data chunk. Returns the list of FLAC frames."""
        a = np.ascontiguousarray(np.frombuffer(wav_bytes, np.uint8) if not isinstance(wav_bytes, np.ndarray) else wav_bytes, dtype=np.uint8)
        bpf = self.channels * (self.bits // 8)
        total = a.size // bpf
        frames = []
        per = self.block_size * self.max_blocks
        out = np.empty(self.max_frame_bytes * self.max_blocks, np.uint8)
        off = (C.c_size_t * self.max_blocks)()
        ln = (C.c_size_t * self.max_blocks)()
        nf = C.c_int32(0)
        done = 0
        while done < total:
            n = min(per, total - done)
            chunk = a[done * bpf:(done + n) * bpf]
            rc = self._L.b200_flac_encode_host(self._h, chunk.ctypes.data, n, self._frames, out.ctypes.data, out.size, off, ln, C.byref(nf))
            if rc:
                raise B200Error(rc, self._L.b200_last_error().decode())
            frames += [out[off[i]:off[i] + ln[i]].tobytes() for i in range(nf.value)]
            self._frames += nf.value
            done += n
        self._samples += total
        return frames

    def codec_private(self, total_samples=None):
        buf = C.create_string_buffer(42)
        self._L.b200_flac_codec_private(self._h, self._samples if total_samples is None else total_samples, buf, 42)
        return buf.raw
