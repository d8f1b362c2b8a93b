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


def run_rawcooked_env(args, cwd, env):
    rc = util.ref_rawcooked()
    if rc is None:
        pytest.skip("oracle/_ref/rawcooked not built")
    p = subprocess.run([rc] + args, cwd=cwd, capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL, env=dict(os.environ, **env))
    return p.returncode, p.stdout + p.stderr


def test_rawcooked_gaps_concat_lists(tmp_path):
    # the reference's gaps.sh (Project/GNU/CLI/test/gaps.sh): numbering gaps make RAWcooked hand the encoder `-f concat`
    # file lists instead of an image2 pattern, one list per image directory, three video tracks in one package
    name = "gaps1"
    layout, w, h = S.DPX_RGB_16_BE, 64, 48
    numbers = {"image1": [0, 6], "image2": [0, 1], "image3": [0, 1, 5, 6]}
    seed = 400
    for sub, nums in numbers.items():
        d = tmp_path / name / sub
        os.makedirs(d)
        for n in nums:
            seed += 1
            open(d / ("%06d.dpx" % n), "wb").write(S.dpx_file(w, h, layout, S.synth_payload(w, h, layout, seed, "grain"), n))
    code, out = run_rawcooked(["--accept-gaps", "--check", "-b", B200ENC, name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    # without --accept-gaps and with -n the reference refuses before it launches the encoder
    os.remove(tmp_path / (name + ".mkv"))
    code2, out2 = run_rawcooked(["-n", "--check", "-b", B200ENC, name], cwd=str(tmp_path))
    assert OK not in out2


@pytest.mark.parametrize("env", [
    {"B200_FRAMES_IN_FLIGHT": "2"},                                   # 7 frames: four batches through one worker
    {"B200_FRAMES_IN_FLIGHT": "3", "B200_DEVICES": "0,0"},            # 10 frames: four batches over two workers (one GPU, two handles)
    {"B200_FRAMES_IN_FLIGHT": "2", "B200_DEVICES": "0,0,0"},
], ids=["one_worker", "two_workers", "three_workers"])
def test_rawcooked_batches_and_workers(tmp_path, env):
    # the front-end's pipeline: batches in flight, one worker per device entry, packets muxed in frame order
    name = "multi"
    n = 7 if "B200_DEVICES" not in env else 10
    write_dpx_sequence(str(tmp_path / name), n, 320, 240, S.DPX_RGB_16_BE, 2000)
    code, out = run_rawcooked_env(["--check", "-y", "-b", B200ENC, "-slices", "24", name], str(tmp_path), env)
    assert code == 0, out
    assert OK in out, out
    victim = tmp_path / name / ("f_%06d.dpx" % (n - 1))
    b = bytearray(victim.read_bytes())
    b[5000] ^= 0x04
    victim.write_bytes(bytes(b))
    code2, out2 = run_rawcooked(["--check", name + ".mkv", "-o", "./"], cwd=str(tmp_path))
    assert OK not in out2 and ("not same" in out2 or code2 != 0), out2


def test_rawcooked_all_gpus(tmp_path):
    # every visible GPU gets a worker (the default); with one GPU this is the single-worker path again
    import torch
    name = "allgpus"
    n = 4 * max(1, torch.cuda.device_count()) + 1
    write_dpx_sequence(str(tmp_path / name), n, 256, 192, S.DPX_RGB_10_FA_BE, 3000)
    code, out = run_rawcooked_env(["--check", "-y", "-b", B200ENC, name], str(tmp_path), {"B200_FRAMES_IN_FLIGHT": "2"})
    assert code == 0, out
    assert OK in out, out


def test_odd_width_16bit_dpx_through_reference(tmp_path):
    # 16-bit DPX with an odd width: the file's lines carry 2 bytes of padding (the only layout the reference's parser sizes the
    # file by, DPX.cpp:478-482). Whatever the reference's own check makes of such a file, the encoder must read the pixels
    # from the padded rows: the packet it writes equals the oracle's for the padded payload.
    name = "odd"
    w, h, layout = 133, 66, S.DPX_RGB_16_BE
    write_dpx_sequence(str(tmp_path / name), 2, w, h, layout, 4000)
    code, out = run_rawcooked(["-y", "--no-check", "-b", B200ENC, name], cwd=str(tmp_path))
    mkv = tmp_path / (name + ".mkv")
    if not mkv.exists():
        pytest.skip("the reference did not launch the encoder for this input: " + out[-300:])
    data = mkv.read_bytes()
    nh, nv = ffv1_grid(w, h, data)
    want = util.oracle_encode(S.synth_payload(w, h, layout, 4000, "grain"), w, h, layout, nh, nv)
    assert want in data


def ffv1_grid(w, h, mkv_bytes):
    # slice grid of the stream from the `-slices` value the reference chose: try the grids ffmpeg's search can produce
    from rawcooked_b200 import ffv1
    for slices in (4, 6, 9, 12, 16, 20, 24, 25, 30, 36, 42, 49, 56, 64):
        try:
            rec = ffv1.config_record(w, h, S.DPX_RGB_16_BE, slices=slices)
        except Exception:      # noqa: BLE001
            continue
        if rec in mkv_bytes:
            return ffv1.slice_grid(w, h, slices)
    raise AssertionError("no known ConfigurationRecord in the MKV")


@pytest.mark.parametrize("env", [
    {"B200_VERIFY": "1", "B200_FRAMES_IN_FLIGHT": "3"},
    {"B200_VERIFY": "1", "B200_FRAMES_IN_FLIGHT": "2", "B200_DEVICES": "0,0"},
], ids=["one_worker", "two_workers"])
def test_encode_time_verification_on_the_gpu(tmp_path, env):
    # B200_VERIFY=1: every packet is decoded again by the CUDA decoder (include/b200dec.h) and compared with its source payload
    # while the next batch is coded; the reference's own --check then confirms the same thing on the CPU
    name = "verify"
    n = 8
    write_dpx_sequence(str(tmp_path / name), n, 320, 240, S.DPX_RGB_16_BE, 4000)
    code, out = run_rawcooked_env(["--check", "-y", "-b", B200ENC, "-slices", "24", name], str(tmp_path), env)
    assert code == 0, out
    assert OK in out, out
    assert "%d of %d frames decoded again on the GPU" % (n, n) in out, out


def test_rawcooked_output_version_2(tmp_path):
    # `--output-version 2` (Source/CLI/Global.cpp:161-177): the command line carries no -attach (Output.cpp:279-291), the
    # reference appends its reversibility data to the Matroska file after the encoder has returned (Main.cpp:905-929) and then
    # parses the result: the Segment this muxer wrote must be complete and of finite size for that to work
    name = "v2"
    write_dpx_sequence(str(tmp_path / name), 5, 256, 192, S.DPX_RGB_10_FA_BE, 5000)
    code, out = run_rawcooked(["--output-version", "2", "--check", "-y", "-b", B200ENC, "-slices", "4", name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    victim = tmp_path / name / "f_000003.dpx"
    b = bytearray(victim.read_bytes())
    b[6000] ^= 0x08
    victim.write_bytes(bytes(b))
    code2, out2 = run_rawcooked(["--check", name + ".mkv", "-o", "./"], cwd=str(tmp_path))
    assert OK not in out2 and ("not same" in out2 or code2 != 0), out2


def test_framemd5_second_output(tmp_path):
    # `rawcooked --framemd5` adds `-f framemd5 <dir>.framemd5` to the command line (Output.cpp:312-332); the file must equal
    # what FFmpeg's own libraries write for the same frames (tests/golden/framemd5_golden.json), but for the #software line
    import json
    cases = json.load(open(os.path.join(util.ROOT, "tests", "golden", "framemd5_golden.json")))["cases"]
    c = [k for k in cases if k["layout"] == S.DPX_RGB_10_FA_BE][0]
    name = "md5seq"
    d = tmp_path / name
    os.makedirs(d)
    for i in range(c["frames"]):
        payload = S.synth_payload(c["w"], c["h"], c["layout"], c["seed"] + i)
        open(d / ("f_%06d.dpx" % i), "wb").write(S.dpx_file(c["w"], c["h"], c["layout"], payload, i))
    code, out = run_rawcooked(["--framemd5", "-framerate", "30000/1001", "--check", "-y", "-b", B200ENC, "-slices", "4", name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    got = (tmp_path / (name + ".framemd5")).read_text()
    strip = lambda t: "\n".join(l for l in t.splitlines() if not l.startswith("#software:"))
    assert strip(got) == strip(c["text"]), got
