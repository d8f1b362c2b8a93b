"""End-to-end drop-in test (GPU): the UNMODIFIED reference CLI (oracle/_ref/rawcooked) analyses the inputs, writes its
reversibility sidecar, launches OUR encoder through --bin-name exactly where it would launch ffmpeg
(/root/reference/Source/CLI/Output.cpp:356), then runs its own `--check` on the MKV we wrote. The test passes when the
reference prints "Reversibility was checked, no issue detected." — the same success string its test suite greps for
(Project/GNU/CLI/test/test2.sh:37).

Two reference flows are used: `rawcooked --check -y -b <bin> <dir>` (encode + full check) and `rawcooked --all ...`.
`--all` additionally switches on the reference's output-MD5 pass, which keeps a pointer into the first mmap of the MKV
(`output_hash::Buffer`, Source/Lib/Compressed/Matroska/Matroska.cpp:79-98) across `filemap::Remap`
(Source/Lib/Utils/FileIO/FileIO.cpp:258-285, munmap + mmap after every MiB parsed). For an MKV between 1 MiB and a few
MiB the new mapping lands elsewhere and the reference itself segfaults — with any encoder, ours or ffmpeg (reproduced with
a shell script standing in for the encoder). `--all` is therefore exercised on a package whose MKV stays below 1 MiB."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import util
from rawcooked_b200 import synth as S

pytestmark = pytest.mark.gpu
B200ENC = os.path.join(util.ROOT, "rawcooked_b200", "b200enc")
OK = "Reversibility was checked, no issue detected."


def run_rawcooked(args, cwd):
    rc = util.ref_rawcooked()
    if rc is None:
        pytest.skip("oracle/_ref/rawcooked not built")
    p = subprocess.run([rc] + args, cwd=cwd, capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
    return p.returncode, p.stdout + p.stderr


def write_dpx_sequence(d, n, w, h, layout, seed0):
    os.makedirs(d, exist_ok=True)
    for i in range(n):
        payload = S.synth_payload(w, h, layout, seed0 + i, "grain")
        open(os.path.join(d, "f_%06d.dpx" % i), "wb").write(S.dpx_file(w, h, layout, payload, i))


@pytest.mark.parametrize("name,mode,w,h,layout,n,extra", [
    ("config1_8bit", "--check", 640, 480, S.DPX_RGB_8, 10, []),               # BASELINE config 1
    ("rgb10", "--all", 256, 192, S.DPX_RGB_10_FA_BE, 4, ["-slices", "4"]),
    ("rgb12packed", "--all", 264, 100, S.DPX_RGB_12_PACKED_BE, 3, []),
    ("rgb16", "--check", 320, 240, S.DPX_RGB_16_BE, 5, ["-slices", "24"]),
])
def test_rawcooked_with_b200enc(tmp_path, name, mode, w, h, layout, n, extra):
    seq = tmp_path / name
    write_dpx_sequence(str(seq), n, w, h, layout, 1000)
    code, out = run_rawcooked([mode, "-y", "-b", B200ENC] + extra + [name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    mkv = tmp_path / (name + ".mkv")
    assert mkv.exists() and mkv.stat().st_size > 1000
    # the check must be a real one: flip one payload byte in a source frame and re-check the MKV against the files
    victim = seq / "f_000002.dpx"
    b = bytearray(victim.read_bytes())
    b[4000] ^= 0x10
    victim.write_bytes(bytes(b))
    code2, out2 = run_rawcooked(["--check", name + ".mkv", "-o", "./"], cwd=str(tmp_path))
    assert OK not in out2 and ("not same" in out2 or code2 != 0), out2


def test_rawcooked_dpx_plus_wav(tmp_path):
    # video + audio in one package: FFV1 + FLAC together (BASELINE config 4 in small)
    name = "pkg"
    seq = tmp_path / name
    write_dpx_sequence(str(seq), 6, 320, 240, S.DPX_RGB_16_BE, 50)
    pcm = S.wav_pcm(2, 48000, 24, 12000, 77)
    open(seq / "audio.wav", "wb").write(S.wav_file(pcm, 48000, 24))
    code, out = run_rawcooked(["--check", "-y", "-b", B200ENC, name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out


def test_rawcooked_tiff_plus_6ch_96k_wav(tmp_path):
    # BASELINE config 4 in small: 16-bit RGB TIFF sequence + 24-bit / 96 kHz / 6-channel WAV (FFV1 + FLAC together)
    name = "tiffpkg"
    seq = tmp_path / name
    os.makedirs(seq)
    w, h, layout = 192, 108, S.TIFF_RGB_16_LE
    for i in range(5):
        payload = S.synth_payload(w, h, layout, 300 + i, "grain")
        open(seq / ("t_%06d.tif" % i), "wb").write(S.tiff_file(w, h, layout, payload))
    pcm = S.wav_pcm(6, 96000, 24, 20000, 78)
    open(seq / "audio.wav", "wb").write(S.wav_file(pcm, 96000, 24))
    code, out = run_rawcooked(["--check", "-y", "-b", B200ENC, name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    # flip one audio byte: the check against the source files must now fail
    victim = seq / "audio.wav"
    b = bytearray(victim.read_bytes())
    b[len(b) // 2] ^= 0x01
    victim.write_bytes(bytes(b))
    code2, out2 = run_rawcooked(["--check", name + ".mkv", "-o", "./"], cwd=str(tmp_path))
    assert OK not in out2 and ("not same" in out2 or code2 != 0), out2
