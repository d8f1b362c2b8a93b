"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, must be byte-identical to
  * FFmpeg's own bitstream on the committed golden vectors (tests/golden/ffv1_golden.npz),
  * the oracle (oracle/ffv1_oracle.c) on seeded inputs of every layout, ragged grids, edge content,
and every packet must decode through the UNMODIFIED reference decoder (oracle/_ref) to the input bytes."""
import os

import numpy as np
import pytest

import util
from rawcooked_b200 import ffv1, synth as S

pytestmark = pytest.mark.gpu



@pytest.mark.parametrize("i", range(util.golden_count()))
def test_cuda_matches_ffmpeg_golden(i):
    w, h, layout, slices, context, ec, payload, rec, pkt = util.golden_case(i)
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, context=context, slicecrc=ec, max_frames=2)
    try:
        assert enc.config_record == rec
        out = enc.encode([payload, payload])
        assert out[0] == pkt
        assert out[1] == pkt
    finally:
        enc.close()


@pytest.mark.parametrize("layout", sorted(S.LAYOUT_BITS))
@pytest.mark.parametrize("w,h,slices", [(96, 64, 4), (200, 150, 6), (131, 77, 9)])
def test_cuda_matches_oracle_and_reference_decoder(layout, w, h, slices):
    for context in (1, 0):
        enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, context=context, max_frames=4)
        try:
            nh, nv = enc.grid
            frames = [S.synth_payload(w, h, layout, 40 + k, kind) for k, kind in enumerate(("grain", "flat", "white", "const"))]
            pkts = enc.encode(frames)
            assert enc.config_record == util.oracle_record(w, h, layout, nh, nv, context)
            for f, p in zip(frames, pkts):
                assert p == util.oracle_encode(f, w, h, layout, nh, nv, context)
                if util.ref_available():
                    assert util.ref_decode(enc.config_record, p, w, h, layout) == f.tobytes()
        finally:
            enc.close()


def test_band_boundaries_and_many_frames():
    # slice heights that are not multiples of the band height, more frames than one warp of slices
    w, h, layout, slices = 160, 203, S.DPX_RGB_16_BE, 6
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=12)
    try:
        nh, nv = enc.grid
        frames = [S.synth_payload(w, h, layout, 900 + k, "grain" if k % 3 else "flat") for k in range(12)]
        pkts = enc.encode(frames)
        for f, p in zip(frames, pkts):
            assert p == util.oracle_encode(f, w, h, layout, nh, nv)
        # fewer frames than max_frames, handle reuse
        pkts2 = enc.encode(frames[3:8])
        assert pkts2 == pkts[3:8]
    finally:
        enc.close()


def test_black_bars_and_black_frames():
    # runs of zero residuals inside grainy rows (pillarbox / letterbox mattes) and whole black frames: the zero-run path of
    # k_model (one_state iterated through its power tables) next to the per-sample paths
    w, h, layout, slices = 400, 120, S.DPX_RGB_16_BE, 4
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=3)
    try:
        nh, nv = enc.grid
        g = S.synth_payload(w, h, layout, 77, "grain").reshape(h, w, 6).copy()
        g[:, : w // 5] = 0
        g[:, 4 * w // 5:] = 0
        g[: h // 6] = 0
        g[h // 2, w // 5: w // 2] = 0x40                 # a constant non-black run inside a row
        frames = [g.reshape(-1), np.zeros(w * h * 6, dtype=np.uint8), np.full(w * h * 6, 0xFF, dtype=np.uint8)]
        pkts = enc.encode(frames)
        for f, p in zip(frames, pkts):
            assert p == util.oracle_encode(f, w, h, layout, nh, nv)
            if util.ref_available():
                assert util.ref_decode(enc.config_record, p, w, h, layout) == f.tobytes()
    finally:
        enc.close()


def test_config2_frame_2k_10bit():
    # one frame of BASELINE config 2 (2048x1556 10-bit Filled-A BE, -slices 4): wide slices (1024 px)
    w, h, layout = 2048, 1556, S.DPX_RGB_10_FA_BE
    enc = ffv1.FFV1Encoder(w, h, layout, slices=4, max_frames=1)
    try:
        f = S.synth_payload(w, h, layout, 2000)
        p = enc.encode([f])[0]
        assert p == util.oracle_encode(f, w, h, layout, 2, 2)
        if util.ref_available():
            assert util.ref_decode(enc.config_record, p, w, h, layout, threads=4) == f.tobytes()
    finally:
        enc.close()


def test_config3_frame_4k_16bit_roundtrip():
    # one frame of BASELINE config 3 (3840x2160 16-bit BE, -slices 24): oracle equality + reference decode
    w, h, layout = 3840, 2160, S.DPX_RGB_16_BE
    enc = ffv1.FFV1Encoder(w, h, layout, slices=24, max_frames=2)
    try:
        assert enc.grid == (6, 4)
        frames = [S.synth_payload(w, h, layout, 3000 + k) for k in range(2)]
        pkts = enc.encode(frames)
        assert pkts[0] == util.oracle_encode(frames[0], w, h, layout, 6, 4)
        if util.ref_available():
            assert util.ref_decode(enc.config_record, pkts[1], w, h, layout, threads=8) == frames[1].tobytes()
    finally:
        enc.close()


def test_config5_frame_8k_12bit():
    # one frame of BASELINE config 5 (7680x4320 12-bit Filled-A BE, -slices 64 = 8x8): 960-px slices, column segments when a
    # row's records outgrow the stage
    w, h, layout = 7680, 4320, S.DPX_RGB_12_FA_BE
    enc = ffv1.FFV1Encoder(w, h, layout, slices=64, max_frames=1)
    try:
        assert enc.grid == (8, 8)
        f = S.synth_payload(w, h, layout, 5000)
        p = enc.encode([f])[0]
        assert p == util.oracle_encode(f, w, h, layout, 8, 8)
        if util.ref_available():
            assert util.ref_decode(enc.config_record, p, w, h, layout, threads=8) == f.tobytes()
    finally:
        enc.close()


def test_two_host_batches_in_flight():
    # b200_ffv1_submit_host: batch i+1 is coded while the packets of batch i are fetched; fetches come back oldest first
    w, h, layout, slices = 256, 144, S.DPX_RGB_10_FA_BE, 4
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=3)
    try:
        nh, nv = enc.grid
        batches = [[S.synth_payload(w, h, layout, 500 + 10 * b + k, "grain" if (b + k) % 2 else "flat") for k in range(3 - (b == 2))] for b in range(4)]
        want = [[util.oracle_encode(f, w, h, layout, nh, nv) for f in bt] for bt in batches]
        enc.submit(batches[0])
        enc.submit(batches[1])
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[0]))] == want[0]
        enc.submit(batches[2])
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[1]))] == want[1]
        enc.submit(batches[3])
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[2]))] == want[2]
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[3]))] == want[3]
        # and the synchronous call still works afterwards
        assert enc.encode(batches[1]) == want[1]
    finally:
        enc.close()


def test_prefetch_ahead_of_the_fetch():
    # b200_ffv1_prefetch_host: the upload of batch i+1 starts before the packets of batch i-1 are fetched; the following submit
    # only launches kernels. Also: a prefetch that the submit does not match (other frames) is ignored, not used
    w, h, layout, slices = 200, 150, S.DPX_RGB_16_BE, 6
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=3)
    try:
        nh, nv = enc.grid
        batches = [[S.synth_payload(w, h, layout, 800 + 10 * b + k, "grain" if (b + k) % 2 else "white") for k in range(3 - (b % 2))] for b in range(6)]
        want = [[util.oracle_encode(f, w, h, layout, nh, nv) for f in bt] for bt in batches]
        enc.prefetch(batches[0])                     # nothing in flight: does nothing
        enc.submit(batches[0])
        enc.submit(batches[1])
        for b in range(2, 6):
            enc.prefetch(batches[b] if b != 4 else batches[0])      # batch 4: a prefetch of the wrong frames
            assert [p.tobytes() for p in enc.fetch_packets(len(batches[b - 2]))] == want[b - 2]
            enc.submit(batches[b])
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[4]))] == want[4]
        assert [p.tobytes() for p in enc.fetch_packets(len(batches[5]))] == want[5]
        assert enc.encode(batches[2]) == want[2]
    finally:
        enc.close()


def test_device_resident_entry_point():
    import torch
    w, h, layout, slices = 320, 240, S.DPX_RGB_16_BE, 4
    n = 5
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=n)
    try:
        frames = [S.synth_payload(w, h, layout, 70 + k) for k in range(n)]
        d = torch.from_numpy(np.concatenate(frames)).cuda()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            enc.encode_device(d.data_ptr(), n, stream.cuda_stream)
        stream.synchronize()
        arena, off, ln = enc.packets_device(n)
        assert arena and off[0] == 0 and all(off[i + 1] == off[i] + ln[i] for i in range(n - 1))
        pkts = enc.fetch_packets(n)
        nh, nv = enc.grid
        for f, p in zip(frames, pkts):
            assert p.tobytes() == util.oracle_encode(f, w, h, layout, nh, nv)
        st = enc.stats()
        assert st["launches"] > 0 and st["bins"] > 0
    finally:
        enc.close()


# ---------------------------------------------------------------------------------------------------------------------
# more (frame, slice, plane-set) items than k_model has CTAs: the persistent grid's work loop, all three kernel variants
# (k_model: > 8 bit with the bank-replicated one_state table; k_model_compact: 8 bit, 28-byte state rows; the lean variant is
# what wide slices fall back to) and every way the pipeline can be scheduled (SM partition on/off, where k_emit runs,
# three band-buffer sets)
MANY = [
    (S.DPX_RGB_16_BE, 320, 240, 24, 16),     # 16 x 24 x 2 = 768 items
    (S.DPX_RGB_10_FA_BE, 320, 240, 24, 16),
    (S.DPX_RGB_8, 320, 240, 24, 16),         # k_model_compact
    (S.DPX_RGB_12_PACKED_BE, 328, 120, 12, 20),
]


def _many(layout, w, h, slices, n, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, max_frames=n)
        try:
            nh, nv = enc.grid
            assert n * nh * nv * 2 > 148 * 2
            kinds = ("grain", "flat", "grain", "white", "const")
            frames = [S.synth_payload(w, h, layout, 7000 + k, kinds[k % len(kinds)]) for k in range(n)]
            pkts = enc.encode(frames)
            for k, (f, p) in enumerate(zip(frames, pkts)):
                assert p == util.oracle_encode(f, w, h, layout, nh, nv), "frame %d differs from the oracle" % k
            # second call on the same handle (scratch re-use), fewer frames
            again = enc.encode(frames[5:9])
            assert again == pkts[5:9]
            if util.ref_available():
                assert util.ref_decode(enc.config_record, pkts[-1], w, h, layout) == frames[-1].tobytes()
        finally:
            enc.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("layout,w,h,slices,n", MANY)
def test_more_items_than_ctas(layout, w, h, slices, n):
    _many(layout, w, h, slices, n)


@pytest.mark.parametrize("env", [
    {"B200_NO_PARTITION": "1"},
    {"B200_NO_PARTITION": "1", "B200_KPAR": "3", "B200_MODEL_RESERVE": "36"},
    {"B200_RANGE_SMS": "8"},
    {"B200_RANGE_SMS": "24", "B200_KPAR": "3"},
    {"B200_RANGE_SMS": "16", "B200_EMIT_SMS": "16"},
    {"B200_RANGE_SMS": "32", "B200_EMIT_MODE": "2"},
], ids=lambda e: ",".join("%s=%s" % (k[5:], v) for k, v in e.items()))
def test_scheduling_variants_bit_exact(env):
    _many(S.DPX_RGB_16_BE, 320, 240, 24, 16, env)
    _many(S.DPX_RGB_8, 256, 120, 12, 14, env)


def test_slicecrc0_matches_oracle_and_reference():
    # -slicecrc 0: ec = 0 in the ConfigurationRecord, 3-byte slice tails (FFV1_Slice.cpp:247-253 reads the CRC only when ec)
    w, h, layout, slices = 200, 150, S.DPX_RGB_16_BE, 6
    enc = ffv1.FFV1Encoder(w, h, layout, slices=slices, slicecrc=0, max_frames=3)
    try:
        nh, nv = enc.grid
        assert enc.config_record == util.oracle_record(w, h, layout, nh, nv, 1, 0)
        frames = [S.synth_payload(w, h, layout, 60 + k) for k in range(3)]
        for f, p in zip(frames, enc.encode(frames)):
            assert p == util.oracle_encode(f, w, h, layout, nh, nv, 1, 0)
            if util.ref_available():
                assert util.ref_decode(enc.config_record, p, w, h, layout) == f.tobytes()
    finally:
        enc.close()


def test_odd_width_dpx_rows():
    # 8-bit and 16-bit DPX pad every line to 32 bits (DPX.cpp:478-482): odd widths shift every row by the padding
    for layout, w, h in ((S.DPX_RGB_16_BE, 133, 66), (S.DPX_RGB_16_LE, 67, 45), (S.DPX_RGB_8, 65, 47), (S.DPX_RGB_8, 130, 61)):
        enc = ffv1.FFV1Encoder(w, h, layout, slices=4, max_frames=2)
        try:
            nh, nv = enc.grid
            assert enc.frame_bytes == S.frame_bytes(w, h, layout)
            frames = [S.synth_payload(w, h, layout, 80 + k) for k in range(2)]
            for f, p in zip(frames, enc.encode(frames)):
                assert p == util.oracle_encode(f, w, h, layout, nh, nv)
                if util.ref_available():
                    assert util.ref_decode(enc.config_record, p, w, h, layout) == f.tobytes()
        finally:
            enc.close()
