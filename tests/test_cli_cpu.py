"""Drop-in boundary on the CPU box (no GPU): the parts of the front-end that need no device — the argv grammar RAWcooked
emits (/root/reference/Source/CLI/Output.cpp:81-332), the WAV payload locator, the Matroska writer with the reversibility
attachment ahead of the first Cluster — exercised by the UNMODIFIED reference CLI with `-c:a copy` on an audio-only
package, then verified by the reference's own `--check`. The video / FLAC paths need the device and live in
test_cli_gpu.py; here the front-end must also refuse, loudly, to encode video without one."""
import os
import subprocess

import numpy as np
import pytest

import util
from rawcooked_b200 import synth as S

B200ENC = os.path.join(util.ROOT, "rawcooked_b200", "b200enc")
OK = "Reversibility was checked, no issue detected."
needs = pytest.mark.skipif(util.ref_rawcooked() is None or not os.path.exists(B200ENC), reason="oracle/_ref/rawcooked or b200enc not built")


def run_rawcooked(args, cwd):
    p = subprocess.run([util.ref_rawcooked()] + args, cwd=cwd, capture_output=True, text=True, timeout=300, stdin=subprocess.DEVNULL)
    return p.returncode, p.stdout + p.stderr


@needs
@pytest.mark.parametrize("ch,rate,bits,n", [(2, 48000, 16, 30000), (6, 96000, 24, 50000)])
def test_audio_only_package_pcm_copy(tmp_path, ch, rate, bits, n):
    name = "aud"
    d = tmp_path / name
    os.makedirs(d)
    open(d / "a.wav", "wb").write(S.wav_file(S.wav_pcm(ch, rate, bits, n, 5), rate, bits))
    code, out = run_rawcooked(["--check", "-y", "-c:a", "copy", "-b", B200ENC, name], cwd=str(tmp_path))
    assert code == 0, out
    assert OK in out, out
    mkv = (tmp_path / (name + ".mkv")).read_bytes()
    assert mkv[:4] == b"\x1a\x45\xdf\xa3" and b"A_PCM/INT/LIT" in mkv and b"RAWcooked reversibility data" in mkv
    # a real check: one flipped source byte must be reported
    b = bytearray((d / "a.wav").read_bytes())
    b[len(b) // 2] ^= 0x01
    (d / "a.wav").write_bytes(bytes(b))
    code2, out2 = run_rawcooked(["--check", name + ".mkv", "-o", "./"], cwd=str(tmp_path))
    assert OK not in out2 and ("not same" in out2 or code2 != 0), out2


@needs
def test_video_needs_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    name = "vid"
    d = tmp_path / name
    os.makedirs(d)
    for i in range(2):
        open(d / ("f_%06d.dpx" % i), "wb").write(S.dpx_file(64, 48, S.DPX_RGB_16_BE, S.synth_payload(64, 48, S.DPX_RGB_16_BE, i), i))
    code, out = run_rawcooked(["--check", "-y", "-b", B200ENC, name], cwd=str(tmp_path))
    assert code != 0 and "no CUDA device" in out and OK not in out, out


@needs
def test_output_grammar_of_the_front_end(tmp_path):
    # what follows the Matroska output on the command line (Output.cpp:306-332): one `-f framemd5 <file>` is accepted, anything
    # else is refused with a message and a non-zero status, before any device is touched
    wav = tmp_path / "a.wav"
    wav.write_bytes(S.wav_file(S.wav_pcm(2, 48000, 16, 4800, 5), 48000, 16))
    base = [B200ENC, "-xerror", "-i", str(wav), "-c:a", "copy", "-c:v", "ffv1", "-coder", "1", "-context", "1", "-f", "matroska", "-g", "1",
            "-level", "3", "-slicecrc", "1", "-y", "-f", "matroska", str(tmp_path / "o.mkv")]
    p = subprocess.run(base + ["-an", "-f", "framemd5", str(tmp_path / "o.framemd5")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr                                    # audio only: nothing to hash, the MKV is written
    assert (tmp_path / "o.mkv").read_bytes()[:4] == b"\x1a\x45\xdf\xa3"
    p = subprocess.run(base + ["-f", "md5", str(tmp_path / "o.md5")], capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "framemd5" in p.stderr
    p = subprocess.run(base + ["-f", "framemd5", str(tmp_path / "1.framemd5"), "-f", "framemd5", str(tmp_path / "2.framemd5")],
                       capture_output=True, text=True, timeout=60)
    assert p.returncode != 0
    # -n with an existing output: refuse like ffmpeg does (Output.cpp:85-87)
    p = subprocess.run([a if a != "-y" else "-n" for a in base], capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "already exists" in p.stderr
    # option values RAWcooked never emits are refused, not ignored
    p = subprocess.run([a if a != "1" or i != base.index("-coder") + 1 else "0" for i, a in enumerate(base)], capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "-coder" in p.stderr
