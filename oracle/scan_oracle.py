"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the padding-bit test of the reference's DPX parser,
dpx::ParseBuffer, /root/reference/Source/Lib/Uncompressed/DPX/DPX.cpp:500-608, for the RGB flavors 0..7
(DPX_Tested, DPX.cpp:184-193). Returns what the reference derives from a payload: how many tested units carry a non-zero
padding bit, the offset of the first one (In_FirstNonZero, :577-586) and the `In` buffer (:587-599: the payload with
everything but the tested padding bits cleared). The reference's own MD5 is reached through oracle/_ref (ref_md5)."""
import numpy as np

BITS = {0: 8, 1: 10, 2: 10, 3: 12, 4: 12, 5: 12, 6: 16, 7: 16}
FILLED = {1: ("<u4", 0x3, 0), 2: (">u4", 0x3, 3), 3: ("<u2", 0xF, 0), 5: (">u2", 0xF, 1)}   # dtype, mask on the value, tested byte


def row_bytes(width, layout):
    bits = BITS[layout]
    if layout in (1, 2):
        return 4 * width
    if layout in (3, 5):
        return 6 * width
    return ((width * bits * 3 + 31) // 32) * 4          # Packed: rows end on a 32-bit boundary (DPX.cpp:478-482)


def padding_test(payload, width, height, layout):
    """-> (nonzero units, first offset or None, masked payload as uint8 array)"""
    p = np.frombuffer(bytes(payload), np.uint8)
    out = np.zeros_like(p)
    if layout in FILLED:
        # Filled A (DPX.cpp:523-534): Step = 4 (10 bit) or 2 (12 bit), the tested byte is the one holding the low bits
        dt, mask, byte = FILLED[layout]
        step = np.dtype(dt).itemsize
        tested = p[byte::step] & mask                    # Buffer[i] & Mask, i += Step
        out[byte::step] = tested
        nz = np.flatnonzero(tested)
        return int(nz.size), (int(nz[0]) * step + byte if nz.size else None), out
    # Packed (DPX.cpp:507-521): only the last word of a row can hold padding
    used = width * BITS[layout] * 3
    rem = used % 32
    if not rem:
        return 0, None, out
    rb = row_bytes(width, layout)
    eol = (used // 32) * 4
    assert eol + 4 == rb
    rows = p.reshape(height, rb)
    words = rows[:, eol:eol + 4].copy().view(">u4")[:, 0].astype(np.uint64)     # ntoh(*(uint32_t*)(Buffer + EOL_i))
    tested = (words & ((0xFFFFFFFF << rem) & 0xFFFFFFFF)).astype(">u4")
    out.reshape(height, rb)[:, eol:eol + 4] = tested.reshape(height, 1).view(np.uint8)
    nz = np.flatnonzero(tested)
    return int(nz.size), (int(nz[0]) * rb + eol if nz.size else None), out


# ---------------------------------------------------------------------------------------------------------------------
# What a `-f framemd5` output hashes (RAWcooked's --framemd5, /root/reference/Source/CLI/Output.cpp:312-332): the frame in
# the pix_fmt FFmpeg's dpx / tiff decoder produces for the flavor, planes back to back, rows without padding.
#   8 bit  -> rgb24          R,G,B bytes
#   10 bit -> gbrp10le       planes G, B, R of 16-bit little-endian samples
#   12 bit -> gbrp12le
#   16 bit -> rgb48le / rgb48be by the byte order of the file
# Pinned to libavcodec's own decoders by tests/golden/framemd5_golden.json (tests/golden/make_framemd5_golden.py).
PIX_FMT = {0: "rgb24", 1: "gbrp10le", 2: "gbrp10le", 3: "gbrp12le", 4: "gbrp12le", 5: "gbrp12le", 6: "rgb48le", 7: "rgb48be",
           32: "rgb24", 33: "rgb48le", 34: "rgb48be"}


def unpack_rgb(payload, width, height, layout):
    """components of every pixel as stored in the file: three (height, width) uint16 arrays R, G, B"""
    p = np.frombuffer(bytes(payload), np.uint8)
    if layout in (0, 32):
        rb = ((3 * width + 3) // 4) * 4 if layout == 0 else 3 * width
        px = p.reshape(height, rb)[:, :3 * width].reshape(height, width, 3).astype(np.uint16)
        return px[..., 0], px[..., 1], px[..., 2]
    if layout in (1, 2):
        v = p.view("<u4" if layout == 1 else ">u4").reshape(height, width).astype(np.uint32)
        return ((v >> 22) & 1023).astype(np.uint16), ((v >> 12) & 1023).astype(np.uint16), ((v >> 2) & 1023).astype(np.uint16)
    if layout in (3, 5):
        v = p.view("<u2" if layout == 3 else ">u2").reshape(height, width, 3) >> 4
        return v[..., 0].astype(np.uint16), v[..., 1].astype(np.uint16), v[..., 2].astype(np.uint16)
    if layout == 4:
        rb = row_bytes(width, 4)
        words = p.reshape(height, rb).view(">u4").astype(np.uint64)
        k = np.arange(3 * width, dtype=np.uint64)
        bit = k * 12
        wi, sh = (bit >> 5).astype(np.int64), bit & 31
        lo = words[:, wi] >> sh
        hi = np.where(sh > 20, words[:, np.minimum(wi + 1, rb // 4 - 1)] << (32 - sh), 0)
        c = ((lo | hi) & 0xFFF).astype(np.uint16).reshape(height, width, 3)
        return c[..., 0], c[..., 1], c[..., 2]
    if layout in (6, 7, 33, 34):
        rb = ((6 * width + 3) // 4) * 4 if layout in (6, 7) else 6 * width
        be = layout in (7, 34)
        v = p.reshape(height, rb)[:, :6 * width].copy().view(">u2" if be else "<u2").reshape(height, width, 3)
        return v[..., 0].astype(np.uint16), v[..., 1].astype(np.uint16), v[..., 2].astype(np.uint16)
    raise ValueError(layout)


def rawvideo_frame(payload, width, height, layout):
    """bytes of the frame as FFmpeg's rawvideo encoder emits them for this flavor (what framemd5 hashes)"""
    R, G, B = unpack_rgb(payload, width, height, layout)
    fmt = PIX_FMT[layout]
    if fmt == "rgb24":
        return np.stack([R, G, B], -1).astype(np.uint8).tobytes()
    if fmt in ("rgb48le", "rgb48be"):
        return np.stack([R, G, B], -1).astype("<u2" if fmt == "rgb48le" else ">u2").tobytes()
    return G.astype("<u2").tobytes() + B.astype("<u2").tobytes() + R.astype("<u2").tobytes()


def framemd5_text(width, height, fps_num, fps_den, sizes_and_digests, software=None):
    """The file libavformat's framehash muxer writes (libavformat/framehash.c ff_framehash_write_header, hashenc.c
    framehash_write_packet), one video stream of rawvideo frames with pts = frame index in a 1/fps time base."""
    out = ["#format: frame checksums", "#version: 2", "#hash: MD5"]
    if software:
        out.append("#software: " + software)
    out += ["#tb 0: %d/%d" % (fps_den, fps_num), "#media_type 0: video", "#codec_id 0: rawvideo",
            "#dimensions 0: %dx%d" % (width, height), "#sar 0: 0/1",
            "#stream#, dts,        pts, duration,     size, hash"]
    for i, (size, dig) in enumerate(sizes_and_digests):
        out.append("%d, %10d, %10d, %8d, %8d, %s" % (0, i, i, 1, size, dig.hex()))
    return "\n".join(out) + "\n"
