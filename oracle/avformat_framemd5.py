"""TEST INFRASTRUCTURE ONLY (oracle/): libavformat's own `framemd5` muxer (62.3.100, bundled with this image's opencv wheel),
driven without headers. RAWcooked's --framemd5 adds `-f framemd5 <file>` as a second output of the ffmpeg command
(/root/reference/Source/CLI/Output.cpp:312-332); this is the file that output produces for a list of raw frames.
Offsets for THIS build: AVFormatContext{pb@32}, AVStream{codecpar@16, time_base@32}, AVCodecParameters{codec_type@0,
codec_id@4, format@44, width@72, height@76}, AVPacket{pts@8, dts@16, data@24, size@32, stream_index@36, duration@64}."""
import ctypes as C
import glob
import os
import tempfile

import avcodec_ffv1 as AV

_avformat = None


def _load():
    global _avformat
    avutil, avcodec = AV._load()
    if _avformat is None:
        for d in AV._LIBDIRS:
            cand = glob.glob(os.path.join(d, "libavformat*.so*"))
            if cand:
                _avformat = C.CDLL(cand[0], mode=C.RTLD_GLOBAL)
                break
        if _avformat is None:
            raise RuntimeError("bundled libavformat not found")
        F = _avformat
        F.avformat_alloc_output_context2.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_char_p, C.c_char_p]
        F.avformat_new_stream.restype = C.c_void_p
        F.avformat_new_stream.argtypes = [C.c_void_p, C.c_void_p]
        F.avio_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        F.avformat_write_header.argtypes = [C.c_void_p, C.c_void_p]
        F.av_write_frame.argtypes = [C.c_void_p, C.c_void_p]
        F.av_write_trailer.argtypes = [C.c_void_p]
        F.avio_closep.argtypes = [C.POINTER(C.c_void_p)]
        F.avformat_free_context.argtypes = [C.c_void_p]
        avcodec.av_packet_alloc.restype = C.c_void_p
        avcodec.av_packet_free.argtypes = [C.c_void_p]
        avutil.av_get_pix_fmt.restype = C.c_int
    return avutil, avcodec, _avformat


def framemd5(frames, width, height, pix_fmt, fps_num, fps_den):
    """The framemd5 text libavformat writes for `frames` (bytes objects: raw frames in pix_fmt), pts = frame index."""
    avutil, avcodec, F = _load()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "out.framemd5").encode()
        ctx = C.c_void_p()
        if F.avformat_alloc_output_context2(C.byref(ctx), None, b"framemd5", path) < 0:
            raise RuntimeError("no framemd5 muxer")
        st = F.avformat_new_stream(ctx, None)
        par = C.cast(st + 16, C.POINTER(C.c_void_p))[0]
        C.cast(par + 0, C.POINTER(C.c_int))[0] = 0           # AVMEDIA_TYPE_VIDEO
        C.cast(par + 4, C.POINTER(C.c_int))[0] = 13          # AV_CODEC_ID_RAWVIDEO
        C.cast(par + 44, C.POINTER(C.c_int))[0] = avutil.av_get_pix_fmt(pix_fmt.encode())
        C.cast(par + 72, C.POINTER(C.c_int))[0] = width
        C.cast(par + 76, C.POINTER(C.c_int))[0] = height
        C.cast(st + 32, C.POINTER(C.c_int))[0] = fps_den     # time base = 1 / frame rate
        C.cast(st + 36, C.POINTER(C.c_int))[0] = fps_num
        pb = C.cast(ctx.value + 32, C.POINTER(C.c_void_p))
        if F.avio_open(pb, path, 2) < 0:
            raise RuntimeError("avio_open failed")
        if F.avformat_write_header(ctx, None) < 0:
            raise RuntimeError("avformat_write_header failed")
        for i, data in enumerate(frames):
            buf = C.create_string_buffer(bytes(data), len(data) + 64)
            pkt = avcodec.av_packet_alloc()
            C.cast(pkt + 8, C.POINTER(C.c_int64))[0] = i
            C.cast(pkt + 16, C.POINTER(C.c_int64))[0] = i
            C.cast(pkt + 24, C.POINTER(C.c_void_p))[0] = C.addressof(buf)
            C.cast(pkt + 32, C.POINTER(C.c_int))[0] = len(data)
            C.cast(pkt + 36, C.POINTER(C.c_int))[0] = 0
            C.cast(pkt + 64, C.POINTER(C.c_int64))[0] = 1
            if F.av_write_frame(ctx, C.c_void_p(pkt)) < 0:
                raise RuntimeError("av_write_frame failed")
            C.cast(pkt + 24, C.POINTER(C.c_void_p))[0] = None
            C.cast(pkt + 32, C.POINTER(C.c_int))[0] = 0
            pp = C.c_void_p(pkt)
            avcodec.av_packet_free(C.byref(pp))
        F.av_write_trailer(ctx)
        F.avio_closep(pb)
        F.avformat_free_context(ctx)
        return open(path.decode()).read()
