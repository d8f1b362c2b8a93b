"""GPU parity tests of the decode + compare path (`--check` side, include/b200dec.h), run with -m gpu on a B200.
The checkers here are the UNMODIFIED reference decoder (oracle/_ref/libref_ffv1dec.so: ffv1_frame::Process + Transform) and the
plain-C restatement of it (oracle/ffv1_oracle.c, second half): the CUDA decoder, called through the C ABI, must produce the same
payload bytes for
  * libavcodec's own packets (tests/golden/ffv1_golden.npz),
  * the CUDA encoder's packets of every layout, ragged grids, both context models, slicecrc 0/1,
and the on-GPU compare must report 0 for the source payload, the exact number of flipped bytes otherwise, and the
reference's error classes for damaged packets."""
import numpy as np
import pytest

import util
from rawcooked_b200 import ffv1, ffv1dec, synth as S

pytestmark = pytest.mark.gpu


def _valid_mask(w, h, layout):
    """bytes of the payload that are pixel data (row padding is not compared)"""
    rb = S.row_bytes(w, layout)
    valid = {S.DPX_RGB_8: 3 * w, S.DPX_RGB_16_LE: 6 * w, S.DPX_RGB_16_BE: 6 * w}.get(layout, rb)
    m = np.zeros((h, rb), bool)
    m[:, :valid] = True
    return m.reshape(-1)


@pytest.mark.parametrize("i", range(util.golden_count()))
def test_decoder_on_ffmpeg_golden_packets(i):
    w, h, layout, slices, context, ec, payload, rec, pkt = util.golden_case(i)
    dec = ffv1dec.FFV1Decoder(w, h, layout, rec, max_frames=2)
    try:
        out, st = dec.decode([pkt, pkt])
        assert st == [0, 0]
        want = util.ref_decode(rec, pkt, w, h, layout) if util.ref_available() else None
        src = np.ascontiguousarray(payload, np.uint8).reshape(-1)
        m = _valid_mask(w, h, layout)
        for o in out:
            o = np.frombuffer(o, np.uint8)
            assert np.array_equal(o[m], src[m])
            if want is not None:
                assert o.tobytes() == want
            assert (o.tobytes(), 0) == util.oracle_decode(rec, pkt, w, h, layout)      # the C restatement of the decoder
        mm, st = dec.check([pkt, pkt], [src, src])
        assert mm == [0, 0] and st == [0, 0]
    finally:
        dec.close()


@pytest.mark.parametrize("layout", sorted(S.LAYOUT_BITS))
@pytest.mark.parametrize("w,h,slices", [(96, 64, 4), (200, 150, 6), (131, 77, 9)])
def test_decoder_inverts_the_encoder_and_equals_reference(layout, w, h, slices):
    for context, ec in ((1, 1), (0, 0)):
        enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, context=context, slicecrc=ec, max_frames=4)
        try:
            frames = [S.synth_payload(w, h, layout, 70 + k, kind) for k, kind in enumerate(("grain", "flat", "white", "const"))]
            pkts = enc.encode(frames)
            rec = enc.config_record
        finally:
            enc.close()
        for spw in (1, 4, 7):
            dec = ffv1dec.FFV1Decoder(w, h, layout, rec, max_frames=4, slices_per_warp=spw)
            try:
                out, st = dec.decode(pkts)
                assert st == [0] * 4
                m = _valid_mask(w, h, layout)
                for f, p, o in zip(frames, pkts, out):
                    o = np.frombuffer(o, np.uint8)
                    assert np.array_equal(o[m], np.asarray(f, np.uint8).reshape(-1)[m])
                    if util.ref_available() and spw == 4:
                        assert o.tobytes() == util.ref_decode(rec, p, w, h, layout)
                    if spw == 1:
                        assert (o.tobytes(), 0) == util.oracle_decode(rec, p, w, h, layout)
                mm, st = dec.check(pkts, frames)
                assert mm == [0] * 4 and st == [0] * 4
            finally:
                dec.close()


def test_compare_counts_flipped_source_bytes_and_flags_damaged_packets():
    w, h, layout, slices = 200, 150, S.DPX_RGB_16_BE, 6
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=3)
    try:
        frames = [S.synth_payload(w, h, layout, 300 + k) for k in range(3)]
        pkts = enc.encode(frames)
        rec = enc.config_record
    finally:
        enc.close()
    dec = ffv1dec.FFV1Decoder(w, h, layout, rec, max_frames=3)
    try:
        # a source file that differs in 3 bytes of frame 1 (what test2.sh does to provoke "files are not same")
        bad = [np.array(f, np.uint8).reshape(-1).copy() for f in frames]
        for pos in (0, 12345, bad[1].size - 1):
            bad[1][pos] ^= 0x40
        mm, st = dec.check(pkts, bad)
        assert mm == [0, 3, 0] and st == [0, 0, 0]
        # a damaged packet: the slice CRC catches it (FFV1-SLICE-slice_crc_parity) and the pixels differ
        dp = [bytearray(p) for p in pkts]
        dp[2][len(dp[2]) // 2] ^= 1
        mm, st = dec.check([bytes(p) for p in dp], frames)
        assert mm[0] == 0 and mm[1] == 0 and st[0] == 0 and st[1] == 0
        assert st[2] & ffv1dec.BAD_CRC and mm[2] > 0
        # a truncated packet: the tail walk fails (FFV1_Frame.cpp:172-181)
        mm, st = dec.check([pkts[0], pkts[1][:-5], pkts[2]], frames)
        assert st[0] == 0 and st[2] == 0 and st[1] & ffv1dec.BAD_TAIL
        # not a keyframe: first bin flipped to 0 is refused (intra-only stream)
        if util.ref_available():
            with pytest.raises(RuntimeError):
                util.ref_decode(rec, bytes(dp[2]), w, h, layout)
    finally:
        dec.close()


def test_decoder_many_slices_in_flight_and_device_entry_point():
    # more slices than one wave of warps; packets taken straight from the encoder's device arena, compared with the device
    # input frames (the encode-time verification the CLI runs with B200_VERIFY=1)
    torch = pytest.importorskip("torch")
    w, h, layout, slices, n = 320, 240, S.DPX_RGB_10_FA_BE, 24, 40
    frames = [S.synth_payload(w, h, layout, 500 + k, "grain" if k % 2 else "flat") for k in range(n)]
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=n)
    try:
        fb = enc.frame_bytes
        host = np.concatenate([np.asarray(f, np.uint8).reshape(-1) for f in frames])
        d_in = torch.from_numpy(host).cuda()
        enc.encode_device(d_in.data_ptr(), n)
        torch.cuda.synchronize()
        d_arena, offs, lens = enc.packets_device(n)
        dec = ffv1dec.FFV1Decoder(w, h, layout, enc.config_record, max_frames=n)
        try:
            d_out = torch.zeros(n * fb, dtype=torch.uint8, device="cuda")
            dec.decode_device(d_arena, offs, lens, d_out=d_out.data_ptr(), d_sources=d_in.data_ptr())
            mm, st = dec.result(n)
            assert mm == [0] * n and st == [0] * n
            assert torch.equal(d_out, d_in)
            s = dec.stats()
            assert s["slices"] == n * slices and s["samples"] == n * w * h * 3
            # flip one pixel of the device source: exactly that frame reports it
            d_in[7 * fb + 100] ^= 1
            dec.decode_device(d_arena, offs, lens, d_sources=d_in.data_ptr())
            mm, st = dec.result(n)
            assert mm == [0] * 7 + [1] + [0] * (n - 8)
        finally:
            dec.close()
    finally:
        enc.close()


def test_decoder_config3_frame_4k_16bit():
    w, h, layout, slices = 3840, 2160, S.DPX_RGB_16_BE, 24
    frames = [S.synth_payload(w, h, layout, 4242 + k) for k in range(2)]
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=2)
    try:
        pkts = enc.encode(frames)
        rec = enc.config_record
    finally:
        enc.close()
    dec = ffv1dec.FFV1Decoder(w, h, layout, rec, max_frames=2)
    try:
        out, st = dec.decode(pkts)
        assert st == [0, 0]
        for f, o in zip(frames, out):
            assert o == np.asarray(f, np.uint8).tobytes()
        mm, st = dec.check(pkts, frames)
        assert mm == [0, 0] and st == [0, 0]
    finally:
        dec.close()
