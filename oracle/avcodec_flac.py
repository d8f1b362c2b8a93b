"""TEST/BENCH INFRASTRUCTURE ONLY (oracle/): FFmpeg's own `flac` encoder through the libavcodec 62.11.100 bundled in this image
(see avcodec_ffv1.py). RAWcooked's audio path IS this encoder (`-c:a flac`, /root/reference/Source/CLI/Global.cpp:949-950, default
compression level 5: LPC orders 1..8, 15-bit coefficients). Nothing in the reference pins FLAC bitstreams, so this module is used
for ONE thing: the size of FFmpeg's output on the benchmark signal, next to which the B200 encoder's size is reported
(tests/test_flac.py, bench.py --config 4). Never on the product path.

Struct offsets are for THIS build: AVCodecParameters{codec_type@0, codec_id@4, format@44,
bits_per_raw_sample@60, ch_layout@128, sample_rate@152, frame_size@160}, AVFrame{data@0, nb_samples@112, format@116, pts@136,
ch_layout@384}, AVPacket{size@32}."""
import ctypes as C

import numpy as np

import avcodec_ffv1 as A

AV_CODEC_ID_FLAC = 0x15000 + 12
AV_SAMPLE_FMT_S16, AV_SAMPLE_FMT_S32 = 1, 2


def flac_encoded_bytes(pcm, sample_rate, bits, compression_level=None):
    """Total bytes of the FLAC frames libavcodec produces for int32 pcm [n, channels] (16 or 24 bit)."""
    avutil, avcodec = A._load()
    n, ch = pcm.shape
    codec = avcodec.avcodec_find_encoder_by_name(b"flac")
    if not codec:
        raise RuntimeError("flac encoder not in bundled libavcodec")
    ctx = C.c_void_p(avcodec.avcodec_alloc_context3(C.c_void_p(codec)))
    par = avcodec.avcodec_parameters_alloc()
    avcodec.avcodec_parameters_to_context.argtypes = [C.c_void_p, C.c_void_p]
    avutil.av_channel_layout_default.argtypes = [C.c_void_p, C.c_int]
    avutil.av_channel_layout_copy.argtypes = [C.c_void_p, C.c_void_p]
    avutil.av_frame_unref.argtypes = [C.c_void_p]
    fmt = AV_SAMPLE_FMT_S16 if bits == 16 else AV_SAMPLE_FMT_S32
    C.c_int.from_address(par + 0).value = 1                  # AVMEDIA_TYPE_AUDIO
    C.c_int.from_address(par + 4).value = AV_CODEC_ID_FLAC
    C.c_int.from_address(par + 44).value = fmt
    C.c_int.from_address(par + 60).value = bits
    avutil.av_channel_layout_default(C.c_void_p(par + 128), ch)
    C.c_int.from_address(par + 152).value = sample_rate
    if avcodec.avcodec_parameters_to_context(ctx, C.c_void_p(par)) < 0:
        raise RuntimeError("avcodec_parameters_to_context failed")
    avutil.av_opt_set(ctx, b"time_base", ("1/%d" % sample_rate).encode(), 1)
    if compression_level is not None:
        avutil.av_opt_set(ctx, b"compression_level", str(compression_level).encode(), 1)
    if avcodec.avcodec_open2(ctx, C.c_void_p(codec), None) < 0:
        raise RuntimeError("avcodec_open2(flac) failed")
    avcodec.avcodec_parameters_from_context(C.c_void_p(par), ctx)
    fs = C.c_int.from_address(par + 160).value
    fr = avutil.av_frame_alloc()
    frame = C.c_void_p(fr)
    pkt = C.c_void_p(avcodec.av_packet_alloc())
    total = 0

    def drain():
        nonlocal total
        while avcodec.avcodec_receive_packet(ctx, pkt) >= 0:
            total += C.c_int.from_address(pkt.value + 32).value
            avcodec.av_packet_unref(pkt)
    pts = 0
    for i in range(0, n, fs):
        blk = pcm[i:i + fs]
        avutil.av_frame_unref(frame)
        C.c_int.from_address(fr + 112).value = len(blk)
        C.c_int.from_address(fr + 116).value = fmt
        avutil.av_channel_layout_copy(C.c_void_p(fr + 384), C.c_void_p(par + 128))
        if avutil.av_frame_get_buffer(frame, 0) < 0:
            raise RuntimeError("av_frame_get_buffer failed")
        data = np.ascontiguousarray(blk).astype(np.int16) if bits == 16 else (np.ascontiguousarray(blk).astype(np.int32) << (32 - bits))
        C.memmove(C.c_void_p.from_address(fr).value, data.ctypes.data, data.nbytes)
        C.c_int64.from_address(fr + 136).value = pts
        pts += len(blk)
        if avcodec.avcodec_send_frame(ctx, frame) < 0:
            raise RuntimeError("avcodec_send_frame failed")
        drain()
    avcodec.avcodec_send_frame(ctx, None)
    drain()
    avcodec.avcodec_free_context(C.byref(ctx))
    return total, fs
