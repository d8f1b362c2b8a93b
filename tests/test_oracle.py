"""CPU tests that PIN the oracle (oracle/ffv1_oracle.c) before anything is checked against it:
  * byte-identical to FFmpeg's own bitstream (libavcodec 62.11.100) on the committed golden vectors;
  * every packet decodes through the UNMODIFIED reference decoder (oracle/_ref) to the input bytes —
    the round-trip property all reference tests pin (Project/GNU/CLI/test/test1.sh, test2.sh, slices.sh);
  * CRC and default state-transition table equal the reference's."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from rawcooked_b200 import synth as S

NGOLD = util.golden_count()
needs_ref = pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built (run oracle/build_ref.sh)")


@pytest.mark.parametrize("i", range(NGOLD))
def test_oracle_matches_ffmpeg_golden(i):
    w, h, layout, slices, context, ec, payload, rec, pkt = util.golden_case(i)
    nh, nv = util.oracle_grid(w, h, slices, S.LAYOUT_BITS[layout])
    assert util.oracle_record(w, h, layout, nh, nv, context, ec) == rec
    assert util.oracle_encode(payload, w, h, layout, nh, nv, context, ec) == pkt


@needs_ref
@pytest.mark.parametrize("i", range(NGOLD))
def test_golden_decodes_through_reference(i):
    w, h, layout, slices, context, ec, payload, rec, pkt = util.golden_case(i)
    assert util.ref_decode(rec, pkt, w, h, layout) == payload.tobytes()


@needs_ref
@pytest.mark.parametrize("layout", sorted(S.LAYOUT_BITS))
@pytest.mark.parametrize("w,h,slices", [(48, 36, 4), (70, 50, 6), (33, 31, 4)])
def test_oracle_roundtrip_reference_decoder(layout, w, h, slices):
    for context in (0, 1):
        for kind in ("grain", "white", "flat"):
            payload = S.synth_payload(w, h, layout, 5 + context, kind)
            nh, nv = util.oracle_grid(w, h, slices, S.LAYOUT_BITS[layout])
            rec = util.oracle_record(w, h, layout, nh, nv, context)
            pkt = util.oracle_encode(payload, w, h, layout, nh, nv, context)
            assert util.ref_decode(rec, pkt, w, h, layout) == payload.tobytes()


def test_slice_grid_matches_reference_valid_set():
    # reference test/slices.sh:12 lists the slice counts ffmpeg accepts; spot-check the grids SURVEY.md observed
    assert util.oracle_grid(2048, 1556, 4, 10) == (2, 2)
    assert util.oracle_grid(3840, 2160, 24, 16) == (6, 4)
    assert util.oracle_grid(7680, 4320, 64, 12) == (8, 8)
    assert util.oracle_grid(640, 480, 16, 8) == (4, 4)
    assert util.oracle_grid(3840, 2160, 576, 16) == (32, 18)   # libavcodec 62.11 agrees (not 24x24)
    valid = [4, 6, 9, 12, 15, 16, 20, 24, 25, 28, 30, 35, 36, 42, 49]
    for n in range(4, 50):
        ok = True
        try:
            nh, nv = util.oracle_grid(1920, 1080, n, 10)
            assert nh * nv == n and nv <= nh <= 2 * nv
        except ValueError:
            ok = False
        if n in valid:
            assert ok, n


@needs_ref
def test_crc_and_default_table_equal_reference():
    R = util.ref_decoder()
    O = util.oracle()
    rng = np.random.default_rng(0)
    for n in (1, 7, 64, 1000, 4099):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        crc = O.ffv1o_crc32(d.ctypes.data, n)
        tail = np.array([crc >> 24, (crc >> 16) & 255, (crc >> 8) & 255, crc & 255], np.uint8)
        both = np.concatenate([d, tail])
        assert R.ref_crc32(both.ctypes.data, both.size) == 0      # how the reference checks (FFV1_Slice.cpp:247-249)
    a = (C.c_uint8 * 256)()
    b = (C.c_uint8 * 256)()
    O.ffv1o_default_transitions(a)
    R.ref_default_state_transitions(b)
    assert bytes(a)[8:249] == bytes(b)[8:249]


def test_oracle_bins_consistent():
    w, h, layout = 64, 48, S.DPX_RGB_16_BE
    payload = S.synth_payload(w, h, layout, 9)
    pkt, sizes, bins = util.oracle_encode(payload, w, h, layout, 2, 2, want_sizes=True)
    assert sizes.sum() == len(pkt)
    total = sum(len(util.oracle_slice_bins(payload, w, h, layout, 2, 2, sx, sy)) for sy in range(2) for sx in range(2))
    assert total == bins


# ---------------------------------------------------------------------------------------------------------------------
# decoder restatement (oracle/ffv1_oracle.c, second half): pinned to the unmodified reference decoder and to the sources
@pytest.mark.parametrize("i", range(NGOLD))
def test_oracle_decoder_on_ffmpeg_golden(i):
    w, h, layout, slices, context, ec, payload, rec, pkt = util.golden_case(i)
    got, flags = util.oracle_decode(rec, pkt, w, h, layout)
    assert flags == 0
    if util.ref_available():
        assert got == util.ref_decode(rec, pkt, w, h, layout)
    src = np.ascontiguousarray(payload, np.uint8).reshape(-1)
    rb = S.row_bytes(w, layout)
    valid = {S.DPX_RGB_8: 3 * w, S.DPX_RGB_16_LE: 6 * w, S.DPX_RGB_16_BE: 6 * w}.get(layout, rb)
    a = np.frombuffer(got, np.uint8).reshape(h, rb)[:, :valid]
    assert np.array_equal(a, src.reshape(h, rb)[:, :valid])


@pytest.mark.parametrize("layout", sorted(S.LAYOUT_BITS))
def test_oracle_decoder_inverts_oracle_encoder(layout):
    for (w, h, slices) in ((48, 36, 4), (70, 50, 6), (33, 31, 4)):
        for context, ec in ((1, 1), (0, 0)):
            nh, nv = util.oracle_grid(w, h, slices, S.LAYOUT_BITS[layout])
            rec = util.oracle_record(w, h, layout, nh, nv, context, ec)
            for kind in ("grain", "white", "const"):
                f = S.synth_payload(w, h, layout, 90, kind)
                pkt = util.oracle_encode(f, w, h, layout, nh, nv, context, ec)
                got, flags = util.oracle_decode(rec, pkt, w, h, layout)
                assert flags == 0 and got == np.asarray(f, np.uint8).tobytes(), (layout, w, h, context, ec, kind)


def test_oracle_decoder_flags_damage():
    w, h, layout = 70, 50, S.DPX_RGB_16_BE
    nh, nv = util.oracle_grid(w, h, 6, 16)
    rec = util.oracle_record(w, h, layout, nh, nv)
    f = S.synth_payload(w, h, layout, 91)
    pkt = bytearray(util.oracle_encode(f, w, h, layout, nh, nv))
    bad = bytearray(pkt)
    bad[len(bad) // 2] ^= 1
    got, flags = util.oracle_decode(rec, bytes(bad), w, h, layout)
    assert flags & 2 and got != np.asarray(f, np.uint8).tobytes()          # slice CRC
    got, flags = util.oracle_decode(rec, bytes(pkt[:-5]), w, h, layout)
    assert flags & 1                                                         # tail walk
    bad = bytearray(pkt)
    bad[0] ^= 0x80                                                           # the keyframe bin
    if util.ref_available():
        with pytest.raises(RuntimeError):
            util.ref_decode(rec, bytes(bad), w, h, layout)
    _, flags = util.oracle_decode(rec, bytes(bad), w, h, layout)
    assert flags != 0
