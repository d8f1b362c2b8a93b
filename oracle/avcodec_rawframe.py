"""TEST INFRASTRUCTURE ONLY (oracle/): what FFmpeg hands to a `-f framemd5` output (the second output RAWcooked adds with
--framemd5, /root/reference/Source/CLI/Output.cpp:312-332): the source image decoded by libavcodec's own `dpx` / `tiff`
decoder, in the decoder's pix_fmt, copied plane after plane without row padding (av_image_copy_to_buffer, align 1) — the
bytes the rawvideo encoder emits and the framehash muxer hashes. Driven through the libavcodec 62.11.100 bundled with this
image's opencv wheel (no ffmpeg CLI, no headers); offsets as in oracle/avcodec_ffv1.py."""
import ctypes as C

import avcodec_ffv1 as AV

BPP = {"rgb24": (1, 3), "rgb48be": (1, 6), "rgb48le": (1, 6), "gbrp10le": (3, 2), "gbrp12le": (3, 2), "gbrp16le": (3, 2)}


def decode_image(file_bytes, codec_name):
    """-> (pix_fmt name, width, height, raw frame bytes as the framemd5 muxer sees them)"""
    avutil, avcodec = AV._load()
    avcodec.avcodec_find_decoder_by_name.restype = C.c_void_p
    avcodec.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avutil.av_get_pix_fmt_name.restype = C.c_char_p
    avutil.av_get_pix_fmt_name.argtypes = [C.c_int]
    avutil.av_frame_free.argtypes = [C.c_void_p]
    codec = avcodec.avcodec_find_decoder_by_name(codec_name.encode())
    if not codec:
        raise RuntimeError("decoder %s not in bundled libavcodec" % codec_name)
    ctx = C.c_void_p(avcodec.avcodec_alloc_context3(codec))
    if avcodec.avcodec_open2(ctx, codec, None) < 0:
        raise RuntimeError("avcodec_open2 failed")
    buf = C.create_string_buffer(bytes(file_bytes) + b"\0" * 64, len(file_bytes) + 64)     # AV_INPUT_BUFFER_PADDING_SIZE
    pkt = C.c_void_p(avcodec.av_packet_alloc())
    C.cast(pkt.value + 24, C.POINTER(C.c_void_p))[0] = C.addressof(buf)
    C.cast(pkt.value + 32, C.POINTER(C.c_int))[0] = len(file_bytes)
    frame = C.c_void_p(avutil.av_frame_alloc())
    r = avcodec.avcodec_send_packet(ctx, pkt)
    if r < 0:
        raise RuntimeError("avcodec_send_packet %d" % r)
    r = avcodec.avcodec_receive_frame(ctx, frame)
    if r < 0:
        raise RuntimeError("avcodec_receive_frame %d" % r)
    data = C.cast(frame.value, C.POINTER(C.c_void_p))
    linesize = C.cast(frame.value + 64, C.POINTER(C.c_int))
    width = C.cast(frame.value + 104, C.POINTER(C.c_int))[0]
    height = C.cast(frame.value + 108, C.POINTER(C.c_int))[0]
    fmt = C.cast(frame.value + 116, C.POINTER(C.c_int))[0]
    name = avutil.av_get_pix_fmt_name(fmt).decode()
    planes, bpp = BPP[name]
    raw = bytearray()
    for p in range(planes):
        for y in range(height):
            raw += C.string_at(data[p] + y * linesize[p], width * bpp)
    C.cast(pkt.value + 24, C.POINTER(C.c_void_p))[0] = None
    C.cast(pkt.value + 32, C.POINTER(C.c_int))[0] = 0
    fp = C.c_void_p(frame.value)
    avutil.av_frame_free(C.byref(fp))
    cp = C.c_void_p(ctx.value)
    avcodec.avcodec_free_context(C.byref(cp))
    return name, width, height, bytes(raw)
