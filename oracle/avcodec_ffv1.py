"""TEST/BENCH INFRASTRUCTURE ONLY (oracle/): FFmpeg's own `ffv1` encoder, driven through the
libavcodec 62.11.100 that ships inside this image's opencv wheel (no ffmpeg CLI, no headers).

RAWcooked's encode hot path IS this encoder: the reference shells out to `ffmpeg -c:v ffv1 -coder 1
-context 1 -g 1 -level 3 -slicecrc 1 -slices N` (/root/reference/Source/CLI/Output.cpp:36-378,
defaults at Source/CLI/Global.cpp:938-989).  FFmpeg is an un-vendored, un-pinned third-party
dependency of the reference, so this module is used for two things only:
  * generating the golden packets under tests/golden/ (tests/golden/make_golden.py), which pin the
    C restatement in oracle/ffv1_oracle.c bit-for-bit to FFmpeg's bitstream;
  * the CPU baseline leg of bench.py (`cpu_baseline.kind == "reference"`).
It is never on the product path.

Struct offsets are for THIS build of libavutil 60.8 / libavcodec 62.11 (verified in SURVEY.md §8c):
AVFrame{data@0, linesize@64, width@104, height@108, format@116}, AVPacket{data@24, size@32}.
"""
import ctypes as C
import glob
import os

_LIBDIRS = [
    os.path.join(os.path.dirname(os.__file__), "site-packages", "opencv_python_headless.libs"),
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs",
]

_libs = None


def _load():
    global _libs
    if _libs is not None:
        return _libs
    for d in _LIBDIRS:
        if not os.path.isdir(d):
            continue
        try:
            # no RPATH in these wheels' libs: load everything in the directory with RTLD_GLOBAL,
            # retrying until the dependency order resolves itself
            loaded = {}
            pending = sorted(glob.glob(os.path.join(d, "*.so*")))
            progress = True
            while pending and progress:
                progress = False
                for p in list(pending):
                    try:
                        loaded[os.path.basename(p).split("-")[0]] = C.CDLL(p, mode=C.RTLD_GLOBAL)
                        pending.remove(p)
                        progress = True
                    except OSError:
                        pass
            if "libavcodec" in loaded and "libavutil" in loaded:
                _libs = (loaded["libavutil"], loaded["libavcodec"])
                break
        except OSError:
            continue
    if _libs is None:
        raise RuntimeError("bundled libavcodec not found")
    avutil, avcodec = _libs
    avcodec.avcodec_find_encoder_by_name.restype = C.c_void_p
    avcodec.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avcodec.avcodec_send_frame.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.av_packet_alloc.restype = C.c_void_p
    avcodec.av_packet_unref.argtypes = [C.c_void_p]
    avcodec.avcodec_free_context.argtypes = [C.c_void_p]
    avcodec.avcodec_parameters_alloc.restype = C.c_void_p
    avcodec.avcodec_parameters_from_context.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_version.restype = C.c_uint
    avutil.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    avutil.av_frame_alloc.restype = C.c_void_p
    avutil.av_frame_get_buffer.argtypes = [C.c_void_p, C.c_int]
    avutil.av_frame_make_writable.argtypes = [C.c_void_p]
    avutil.av_get_pix_fmt.argtypes = [C.c_char_p]
    avutil.av_get_pix_fmt.restype = C.c_int
    avutil.av_log_set_level.argtypes = [C.c_int]
    avutil.av_log_set_level(16)
    return _libs


def version():
    _, avcodec = _load()
    v = avcodec.avcodec_version()
    return "%d.%d.%d" % (v >> 16, (v >> 8) & 255, v & 255)


class FFV1Encoder:
    """One libavcodec ffv1 encoder context with RAWcooked's option set."""

    def __init__(self, width, height, pix_fmt, slices, threads=1, context=1, coder=1, level=3, slicecrc=1):
        avutil, avcodec = _load()
        self.avutil, self.avcodec = avutil, avcodec
        codec = avcodec.avcodec_find_encoder_by_name(b"ffv1")
        if not codec:
            raise RuntimeError("ffv1 encoder not in bundled libavcodec")
        ctx = avcodec.avcodec_alloc_context3(codec)
        self.ctx = C.c_void_p(ctx)
        S = 1  # AV_OPT_SEARCH_CHILDREN
        def opt(k, v):
            r = avutil.av_opt_set(self.ctx, k.encode(), str(v).encode(), S)
            if r < 0:
                raise RuntimeError("av_opt_set %s=%s failed %d" % (k, v, r))
        opt("video_size", "%dx%d" % (width, height))
        opt("pixel_format", pix_fmt)
        opt("time_base", "1/24")
        opt("coder", coder)
        opt("context", context)
        opt("g", 1)
        opt("level", level)
        opt("slicecrc", slicecrc)
        if slices:
            opt("slices", slices)
        opt("threads", threads)
        if threads > 1:
            opt("thread_type", "slice")
        r = avcodec.avcodec_open2(self.ctx, codec, None)
        if r < 0:
            raise RuntimeError("avcodec_open2 failed %d" % r)
        par = avcodec.avcodec_parameters_alloc()
        avcodec.avcodec_parameters_from_context(C.c_void_p(par), self.ctx)
        ed = C.c_void_p.from_address(par + 16).value
        eds = C.c_int.from_address(par + 24).value
        self.extradata = C.string_at(ed, eds)
        self.width, self.height = width, height
        self.pix_fmt = pix_fmt
        fr = avutil.av_frame_alloc()
        self.frame = C.c_void_p(fr)
        C.c_int.from_address(fr + 104).value = width
        C.c_int.from_address(fr + 108).value = height
        C.c_int.from_address(fr + 116).value = avutil.av_get_pix_fmt(pix_fmt.encode())
        r = avutil.av_frame_get_buffer(self.frame, 0)
        if r < 0:
            raise RuntimeError("av_frame_get_buffer failed %d" % r)
        self.pkt = C.c_void_p(avcodec.av_packet_alloc())
        self._pts = 0

    def planes(self):
        """(address, linesize) for each data plane of the frame buffer."""
        fr = self.frame.value
        out = []
        for i in range(4):
            p = C.c_void_p.from_address(fr + 8 * i).value
            ls = C.c_int.from_address(fr + 64 + 4 * i).value
            if p:
                out.append((p, ls))
        return out

    def encode_planes(self, plane_arrays):
        """plane_arrays: list of 2-D numpy arrays (rows x row-bytes as the pix_fmt wants them)."""
        import numpy as np
        self.avutil.av_frame_make_writable(self.frame)
        for (addr, ls), arr in zip(self.planes(), plane_arrays):
            a = np.ascontiguousarray(arr)
            rowbytes = a.shape[1] * a.itemsize
            dst = np.ctypeslib.as_array((C.c_uint8 * (ls * self.height)).from_address(addr)).reshape(self.height, ls)
            dst[:, :rowbytes] = a.view(np.uint8).reshape(self.height, rowbytes)
        return self.encode_current()

    def encode_current(self):
        fr = self.frame.value
        C.c_int64.from_address(fr + 136).value = self._pts  # pts (harmless if offset differs: intra-only)
        self._pts += 1
        r = self.avcodec.avcodec_send_frame(self.ctx, self.frame)
        if r < 0:
            raise RuntimeError("send_frame %d" % r)
        r = self.avcodec.avcodec_receive_packet(self.ctx, self.pkt)
        if r < 0:
            raise RuntimeError("receive_packet %d" % r)
        p = self.pkt.value
        data = C.c_void_p.from_address(p + 24).value
        size = C.c_int.from_address(p + 32).value
        out = C.string_at(data, size)
        self.avcodec.av_packet_unref(self.pkt)
        return out

    def close(self):
        if self.ctx:
            self.avcodec.avcodec_free_context(C.byref(self.ctx))
            self.ctx = None
