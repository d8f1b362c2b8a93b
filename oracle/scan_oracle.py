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
