"""Seeded synthetic inputs for tests and bench: image payloads in the exact on-disk
layouts of the DPX/TIFF flavors the reference accepts (Source/Lib/Uncompressed/DPX/DPX.cpp:184-209,
TIFF/TIFF.cpp:157-167; byte layouts per Source/Lib/Transform/Transform.cpp:70-420), full DPX/TIFF/WAV
files around them, and the content models named in the benchmark record (ramps + uniform noise)."""
import struct

import numpy as np

# b200_layout (include/b200enc.h)
DPX_RGB_8, DPX_RGB_10_FA_LE, DPX_RGB_10_FA_BE, DPX_RGB_12_FA_LE = 0, 1, 2, 3
DPX_RGB_12_PACKED_BE, DPX_RGB_12_FA_BE, DPX_RGB_16_LE, DPX_RGB_16_BE = 4, 5, 6, 7
TIFF_RGB_8, TIFF_RGB_16_LE, TIFF_RGB_16_BE = 32, 33, 34

LAYOUT_BITS = {0: 8, 1: 10, 2: 10, 3: 12, 4: 12, 5: 12, 6: 16, 7: 16, 32: 8, 33: 16, 34: 16}
LAYOUT_NAMES = {0: "dpx_rgb8", 1: "dpx_rgb10_fa_le", 2: "dpx_rgb10_fa_be", 3: "dpx_rgb12_fa_le",
                4: "dpx_rgb12_packed_be", 5: "dpx_rgb12_fa_be", 6: "dpx_rgb16_le", 7: "dpx_rgb16_be",
                32: "tiff_rgb8", 33: "tiff_rgb16_le", 34: "tiff_rgb16_be"}


def rgb_content(width, height, bits, seed, kind="grain", noise=None):
    """(R, G, B) uint16 arrays with values < 2**bits.
    kind: 'grain' ramps + uniform noise (default amplitude 2**(bits-8): 16-bit -> +-256, 10-bit -> +-4);
          'flat'  ramps + noise +-4 (16-bit) / +-1; 'white' uniform random; 'zero' all zero;
          'const' a constant colour; 'ramp' noiseless ramps."""
    rng = np.random.default_rng(seed)
    mx = (1 << bits) - 1
    yy, xx = np.mgrid[0:height, 0:width].astype(np.int64)
    if kind == "white":
        return tuple(rng.integers(0, mx + 1, (height, width), dtype=np.int64).astype(np.uint16) for _ in range(3))
    if kind == "zero":
        return tuple(np.zeros((height, width), np.uint16) for _ in range(3))
    if kind == "const":
        return tuple(np.full((height, width), (mx * (k + 1)) // 4, np.uint16) for k in range(3))
    if noise is None:
        noise = {"grain": max(1, 1 << max(bits - 8, 0)), "flat": max(1, 1 << max(bits - 14, 0)), "ramp": 0}[kind]
    out = []
    for k in range(3):
        ramp = (xx * (k + 1) * mx // (3 * max(width - 1, 1)) + yy * (3 - k) * mx // (4 * max(height - 1, 1))) % (mx + 1)
        n = rng.integers(-noise, noise + 1, (height, width), dtype=np.int64) if noise else 0
        out.append(np.clip(ramp + n, 0, mx).astype(np.uint16))
    return tuple(out)


def row_bytes(width, layout):
    if layout == DPX_RGB_8:
        return (width * 3 + 3) & ~3
    if layout == TIFF_RGB_8:
        return width * 3
    if layout in (DPX_RGB_10_FA_LE, DPX_RGB_10_FA_BE):
        return width * 4
    if layout == DPX_RGB_12_PACKED_BE:
        return ((width * 36 + 31) // 32) * 4
    if layout in (DPX_RGB_16_LE, DPX_RGB_16_BE):
        return (width * 6 + 3) & ~3          # DPX lines are padded to 32 bits unless the packing is "Filled" (DPX.cpp:478-482)
    return width * 6


def frame_bytes(width, height, layout):
    return row_bytes(width, layout) * height


def pack_payload(R, G, B, layout):
    """Image payload bytes (numpy uint8, 1-D) of one frame in `layout`."""
    h, w = R.shape
    if layout in (DPX_RGB_8, TIFF_RGB_8):
        rb = row_bytes(w, layout)
        out = np.zeros((h, rb), np.uint8)
        out[:, : 3 * w] = np.stack([R, G, B], -1).astype(np.uint8).reshape(h, 3 * w)
        return out.reshape(-1)
    if layout in (DPX_RGB_10_FA_LE, DPX_RGB_10_FA_BE):
        v = (R.astype(np.uint32) << 22) | (G.astype(np.uint32) << 12) | (B.astype(np.uint32) << 2)
        return v.astype("<u4" if layout == DPX_RGB_10_FA_LE else ">u4").view(np.uint8).reshape(-1)
    if layout in (DPX_RGB_12_FA_LE, DPX_RGB_12_FA_BE, DPX_RGB_16_LE, DPX_RGB_16_BE, TIFF_RGB_16_LE, TIFF_RGB_16_BE):
        sh = 4 if layout in (DPX_RGB_12_FA_LE, DPX_RGB_12_FA_BE) else 0
        le = layout in (DPX_RGB_12_FA_LE, DPX_RGB_16_LE, TIFF_RGB_16_LE)
        v = (np.stack([R, G, B], -1).astype(np.uint16) << sh)
        rows = v.astype("<u2" if le else ">u2").view(np.uint8).reshape(h, 6 * w)
        rb = row_bytes(w, layout)
        if rb != 6 * w:
            rows = np.pad(rows, ((0, 0), (0, rb - 6 * w)))
        return np.ascontiguousarray(rows).reshape(-1)
    if layout == DPX_RGB_12_PACKED_BE:
        comps = np.stack([R, G, B], -1).reshape(h, 3 * w).astype(np.uint64)
        nwords = row_bytes(w, layout) // 4
        bits = ((comps[:, :, None] >> np.arange(12, dtype=np.uint64)) & 1).astype(np.uint8).reshape(h, -1)
        pad = nwords * 32 - bits.shape[1]
        bits = np.pad(bits, ((0, 0), (0, pad))).reshape(h, nwords, 32)
        words = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
        return words.astype(">u4").view(np.uint8).reshape(-1)
    raise ValueError(layout)


def synth_payload(width, height, layout, seed, kind="grain", noise=None):
    R, G, B = rgb_content(width, height, LAYOUT_BITS[layout], seed, kind, noise)
    return pack_payload(R, G, B, layout)


# ---------------------------------------------------------------------------------------------------
# whole files (headers the reference parsers accept: SURVEY.md §8d, DPX.cpp:250-416, TIFF.cpp:380-717,
# WAV.cpp:271-376)

_DPX_DESC = {DPX_RGB_8: (8, 0, ">"), DPX_RGB_10_FA_LE: (10, 1, "<"), DPX_RGB_10_FA_BE: (10, 1, ">"),
             DPX_RGB_12_FA_LE: (12, 1, "<"), DPX_RGB_12_PACKED_BE: (12, 0, ">"), DPX_RGB_12_FA_BE: (12, 1, ">"),
             DPX_RGB_16_LE: (16, 0, "<"), DPX_RGB_16_BE: (16, 0, ">")}


def dpx_file(width, height, layout, payload, frame_number=0, header_size=2048):
    bits, packing, en = _DPX_DESC[layout]
    hdr = bytearray(header_size)
    hdr[0:4] = b"SDPX" if en == ">" else b"XPDS"
    struct.pack_into(en + "I", hdr, 4, header_size)
    hdr[8:12] = b"V2.0"
    struct.pack_into(en + "I", hdr, 16, header_size + len(payload))
    struct.pack_into(en + "I", hdr, 20, 1)
    struct.pack_into(en + "I", hdr, 24, 1664)
    struct.pack_into(en + "I", hdr, 28, 384)
    name = ("frame_%06d" % frame_number).encode()
    hdr[36:36 + len(name)] = name
    struct.pack_into(en + "I", hdr, 660, 0xFFFFFFFF)
    struct.pack_into(en + "H", hdr, 768, 0)
    struct.pack_into(en + "H", hdr, 770, 1)
    struct.pack_into(en + "I", hdr, 772, width)
    struct.pack_into(en + "I", hdr, 776, height)
    struct.pack_into(en + "I", hdr, 780, 0)
    hdr[800] = 50
    hdr[801] = 2
    hdr[802] = 2
    hdr[803] = bits
    struct.pack_into(en + "H", hdr, 804, packing)
    struct.pack_into(en + "H", hdr, 806, 0)
    struct.pack_into(en + "I", hdr, 808, header_size)
    struct.pack_into(en + "I", hdr, 812, 0)
    struct.pack_into(en + "I", hdr, 816, 0)
    return bytes(hdr) + bytes(payload)


def tiff_file(width, height, layout, payload):
    """Baseline TIFF, one strip, RGB, 8 or 16 bit (constraints TIFF.cpp:545-590)."""
    bits = LAYOUT_BITS[layout]
    en = ">" if layout == TIFF_RGB_16_BE else "<"
    entries = []
    data_after_ifd = bytearray()
    n_entries = 10
    ifd_off = 8
    extra_off = ifd_off + 2 + n_entries * 12 + 4
    bps_off = extra_off
    data_after_ifd += struct.pack(en + "HHH", bits, bits, bits)
    strip_off = extra_off + len(data_after_ifd)
    if strip_off & 1:
        data_after_ifd += b"\0"
        strip_off += 1

    def ent(tag, typ, count, value):
        if typ == 3 and count == 1:
            return struct.pack(en + "HHIHH", tag, typ, count, value, 0)
        return struct.pack(en + "HHII", tag, typ, count, value)
    entries.append(ent(256, 4, 1, width))
    entries.append(ent(257, 4, 1, height))
    entries.append(ent(258, 3, 3, bps_off))
    entries.append(ent(259, 3, 1, 1))
    entries.append(ent(262, 3, 1, 2))
    entries.append(ent(273, 4, 1, strip_off))
    entries.append(ent(277, 3, 1, 3))
    entries.append(ent(278, 4, 1, height))
    entries.append(ent(279, 4, 1, len(payload)))
    entries.append(ent(284, 3, 1, 1))
    head = (b"II" if en == "<" else b"MM") + struct.pack(en + "HI", 42, ifd_off)
    ifd = struct.pack(en + "H", n_entries) + b"".join(entries) + struct.pack(en + "I", 0)
    return head + ifd + bytes(data_after_ifd) + bytes(payload)


def wav_pcm(channels, sample_rate, bits, n_samples, seed=77):
    """Interleaved int32 samples (n_samples, channels): per-channel sines + uniform noise +-2^(bits-13)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    amp = (1 << (bits - 1)) * 0.4
    out = np.zeros((n_samples, channels), np.int64)
    nz = max(1, 1 << max(bits - 13, 0))
    for c in range(channels):
        s = amp * np.sin(2 * np.pi * (220.0 * (c + 1)) * t) + amp * 0.3 * np.sin(2 * np.pi * (3300.0 + 17 * c) * t)
        out[:, c] = np.round(s).astype(np.int64) + rng.integers(-nz, nz + 1, n_samples)
    lim = (1 << (bits - 1)) - 1
    return np.clip(out, -lim - 1, lim).astype(np.int32)


def wav_file(pcm, sample_rate, bits):
    """RIFF/WAVE PCM (tag 1), little-endian packed samples."""
    n, ch = pcm.shape
    bps = bits // 8
    if bits == 8:
        raw = (pcm + 128).astype(np.uint8).tobytes()
    else:
        b = pcm.astype("<i4").view(np.uint8).reshape(n, ch, 4)[:, :, :bps]
        raw = np.ascontiguousarray(b).tobytes()
    fmt = struct.pack("<HHIIHH", 1, ch, sample_rate, sample_rate * ch * bps, ch * bps, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(raw)) + raw
    if len(raw) & 1:
        body += b"\0"
    return b"RIFF" + struct.pack("<I", len(body)) + body
