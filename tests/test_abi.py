"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/b200enc.h declares, does the
host-only work (slice grid, payload sizes) correctly, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import util
from rawcooked_b200 import ffv1, synth as S

ROOT = util.ROOT


def declared_symbols():
    syms = set()
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not name.endswith(".h"):
            continue
        hdr = open(os.path.join(ROOT, "include", name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        syms |= set(re.findall(r"\b(b200(?:enc)?_[a-z0-9_]+)\s*\(", hdr))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    L = ffv1.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30 and "b200_ffv1_check_host" in syms and "b200enc_main" in syms
    for s in syms:
        assert hasattr(L, s), "libb200enc.so does not export %s" % s


def test_version_and_error_string():
    L = ffv1.load_library()
    assert L.b200_version() >= (1 << 8)
    assert isinstance(L.b200_last_error(), bytes)


@pytest.mark.parametrize("w,h,n", [(2048, 1556, 4), (3840, 2160, 24), (7680, 4320, 64), (640, 480, 16), (3840, 2160, 576), (1920, 1080, 9)])
def test_slice_grid_equals_oracle(w, h, n):
    assert ffv1.slice_grid(w, h, n) == util.oracle_grid(w, h, n, 16)


def test_slice_grid_rejects_invalid_counts():
    for n in (5, 7, 11, 13):          # not in the set ffmpeg accepts (reference test/slices.sh:12)
        with pytest.raises(ffv1.B200Error):
            ffv1.slice_grid(1920, 1080, n)


@pytest.mark.parametrize("layout", sorted(S.LAYOUT_BITS))
def test_frame_bytes(layout):
    for w, h in ((64, 48), (100, 50), (3840, 2160)):
        assert ffv1.frame_bytes(w, h, layout) == S.frame_bytes(w, h, layout) == util.oracle().ffv1o_frame_bytes(w, h, layout)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ffv1.B200Error) as e:
        ffv1.FFV1Encoder(64, 48, S.DPX_RGB_16_BE, slices=4)
    assert e.value.code == -2          # B200_ERR_NO_DEVICE


def test_product_does_not_touch_oracle():
    # the product path (package + csrc + include) must not reference anything under oracle/
    bad = []
    for base in ("rawcooked_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"oracle[/.]|ffv1o_|libffv1_oracle|libref_", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_host_config_record_equals_ffmpeg_golden():
    # the ConfigurationRecord is host work (quantisation tables, state-transition table, slice grid, CRC): byte-identical to
    # the record libavcodec produced for every golden case, without a device
    import util
    for i in range(util.golden_count()):
        w, h, layout, slices, context, ec, _, rec, _ = util.golden_case(i)
        assert ffv1.config_record(w, h, layout, slices=slices, context=context, slicecrc=ec) == rec, (i, w, h, layout, slices, context, ec)
    with pytest.raises(ffv1.B200Error):
        ffv1.config_record(1920, 1080, S.DPX_RGB_16_BE, slices=7)


def test_config_record_parser_reads_ffmpeg_and_own_records():
    # parameters::Parse on the host (the decoder's side of the ConfigurationRecord): libavcodec's golden records and the
    # encoder's own must read back as the option set that produced them
    import util
    from rawcooked_b200 import ffv1dec
    for i in range(util.golden_count()):
        w, h, layout, slices, context, ec, _, rec, _ = util.golden_case(i)
        nh, nv = ffv1.slice_grid(w, h, slices)
        for r in (rec, ffv1.config_record(w, h, layout, slices=slices, context=context, slicecrc=ec)):
            p = ffv1dec.parse_config_record(r)
            assert (p.version, p.micro_version, p.coder_type, p.colorspace_type) == (3, 4, 2, 1)
            assert p.bits_per_raw_sample == S.LAYOUT_BITS[layout]
            assert (p.chroma_planes, p.alpha_plane, p.log2_h_chroma_subsample, p.log2_v_chroma_subsample) == (1, 0, 0, 0)
            assert (p.num_h_slices, p.num_v_slices, p.ec, p.intra, p.quant_table_set_count, p.crc_ok) == (nh, nv, ec, 1, 2, 1)
            small, large = ((11 ** 3 + 1) // 2, (11 * 11 * 125 + 1) // 2) if p.bits_per_raw_sample == 8 else ((9 ** 3 + 1) // 2, (9 * 9 * 125 + 1) // 2)
            assert list(p.context_count)[:2] == [small, large]
    # a flipped bit is caught by the record's CRC, with the reference's error text
    bad = bytearray(rec)
    bad[5] ^= 4
    with pytest.raises(ffv1.B200Error, match="configuration_record_crc_parity"):
        ffv1dec.parse_config_record(bytes(bad))


def test_decoder_needs_a_device():
    import util
    from rawcooked_b200 import ffv1dec
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    w, h, layout, slices, context, ec, _, rec, _ = util.golden_case(0)
    with pytest.raises(ffv1.B200Error) as e:
        ffv1dec.FFV1Decoder(w, h, layout, rec)
    assert e.value.code == -2


def test_config_record_round_trip_over_the_option_space():
    # host range encoder (ffv1_host.cpp) against host range decoder (ffv1_dec_host.cpp) over layouts x grids x context x ec:
    # what the encoder's record says must be what the decoder's parser reads
    from rawcooked_b200 import ffv1dec
    n = 0
    for layout in sorted(S.LAYOUT_BITS):
        for (w, h) in ((64, 48), (1920, 1080), (4096, 3112), (131, 77)):
            for slices in (4, 6, 9, 12, 16, 24, 30, 64, 576):
                try:
                    nh, nv = ffv1.slice_grid(w, h, slices)
                except ffv1.B200Error:
                    continue
                if nh >= w or nv >= h:
                    continue
                for context in (0, 1):
                    for ec in (0, 1):
                        rec = ffv1.config_record(w, h, layout, slices=slices, context=context, slicecrc=ec)
                        p = ffv1dec.parse_config_record(rec)
                        assert (p.num_h_slices, p.num_v_slices, p.ec, p.bits_per_raw_sample, p.crc_ok) == (nh, nv, ec, S.LAYOUT_BITS[layout], 1)
                        n += 1
    assert n > 500
