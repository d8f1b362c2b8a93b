"""Structure of the Matroska files the front-end writes ((f)1: SeekHead, Cues, finite Segment size), checked with a small
EBML walker written for the test: every SeekHead entry must point at an element with the ID it names, every CuePoint at a
Cluster whose Timestamp is the CueTime, cluster timestamps must not go backwards within a track, SimpleBlock offsets must fit
int16. The element IDs are the ones the reference's parser dispatches on
(/root/reference/Source/Lib/Compressed/Matroska/Matroska.cpp:128-217). Runs on the CPU: an audio-only package with `-c:a copy`
needs no device."""
import os
import subprocess

import pytest

import util
from rawcooked_b200 import synth as S

B200ENC = os.path.join(util.ROOT, "rawcooked_b200", "b200enc")
ID_EBML, ID_SEGMENT, ID_SEEKHEAD, ID_SEEK, ID_SEEKID, ID_SEEKPOS = 0x1A45DFA3, 0x18538067, 0x114D9B74, 0x4DBB, 0x53AB, 0x53AC
ID_INFO, ID_TRACKS, ID_ATTACH, ID_CLUSTER, ID_CUES = 0x1549A966, 0x1654AE6B, 0x1941A469, 0x1F43B675, 0x1C53BB6B
ID_TIMESTAMP, ID_SIMPLEBLOCK, ID_CUEPOINT, ID_CUETIME, ID_CUETRACKPOS, ID_CUETRACK, ID_CUECLUSTERPOS = 0xE7, 0xA3, 0xBB, 0xB3, 0xB7, 0xF7, 0xF1
ID_VOID = 0xEC


def read_id(b, p):
    first = b[p]
    n = 1
    while n <= 4 and not (first & (0x80 >> (n - 1))):
        n += 1
    return int.from_bytes(b[p:p + n], "big"), p + n


def read_size(b, p):
    first = b[p]
    n = 1
    while n <= 8 and not (first & (0x80 >> (n - 1))):
        n += 1
    v = int.from_bytes(b[p:p + n], "big") & ((1 << (7 * n)) - 1)
    unknown = v == (1 << (7 * n)) - 1
    return (None if unknown else v), p + n


def children(b, start, end):
    p = start
    while p < end:
        eid, q = read_id(b, p)
        size, q = read_size(b, q)
        assert size is not None, "element 0x%X of unknown size" % eid
        assert q + size <= end, "element 0x%X runs past its parent" % eid
        yield eid, q, size, p
        p = q + size
    assert p == end


def uint(b, p, n):
    return int.from_bytes(b[p:p + n], "big")


@pytest.mark.skipif(not os.path.exists(B200ENC), reason="b200enc not built")
def test_seekhead_cues_and_clusters(tmp_path):
    wavs = []
    for k, (ch, rate, bits, secs) in enumerate(((2, 48000, 16, 12), (1, 44100, 24, 9))):
        w = tmp_path / ("a%d.wav" % k)
        w.write_bytes(S.wav_file(S.wav_pcm(ch, rate, bits, rate * secs, 5 + k), rate, bits))
        wavs.append(str(w))
    att = tmp_path / "side.bin"
    att.write_bytes(b"reversibility" * 100)
    out = tmp_path / "o.mkv"
    cmd = [B200ENC, "-xerror", "-i", wavs[0], "-i", wavs[1], "-map", "0", "-map", "1", "-c:a", "copy", "-c:v", "ffv1", "-coder", "1", "-context", "1",
           "-f", "matroska", "-g", "1", "-level", "3", "-slicecrc", "1", "-y", "-attach", str(att), "-metadata:s:2", "mimetype=application/octet-stream",
           "-metadata:s:2", "filename=RAWcooked reversibility data", "-f", "matroska", str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    b = out.read_bytes()
    top = list(children(b, 0, len(b)))
    assert [t[0] for t in top] == [ID_EBML, ID_SEGMENT]                 # finite Segment size that ends with the file (Matroska.cpp:494-498)
    _, seg0, seg_size, _ = top[1]
    assert seg0 + seg_size == len(b)
    level1 = list(children(b, seg0, seg0 + seg_size))
    ids = [e[0] for e in level1]
    assert ids[0] == ID_SEEKHEAD
    first_cluster = ids.index(ID_CLUSTER)
    assert ID_ATTACH in ids[:first_cluster] and ID_TRACKS in ids[:first_cluster] and ID_INFO in ids[:first_cluster]   # attachments ahead of the Clusters
    assert ids.count(ID_CUES) == 1 and ids.index(ID_CUES) > first_cluster
    pos_of = {}
    for eid, q, size, p0 in level1:
        pos_of.setdefault(eid, []).append(p0 - seg0)
    # SeekHead: every entry names an element that really starts there
    _, q, size, _ = level1[0]
    seen = set()
    for eid, q2, size2, _ in children(b, q, q + size):
        if eid == ID_VOID:
            continue
        assert eid == ID_SEEK
        sid = spos = None
        for e3, q3, s3, _ in children(b, q2, q2 + size2):
            if e3 == ID_SEEKID:
                sid = uint(b, q3, s3)
            elif e3 == ID_SEEKPOS:
                spos = uint(b, q3, s3)
        assert sid in pos_of and spos in pos_of[sid], "SeekHead entry 0x%X -> %s" % (sid, spos)
        assert read_id(b, seg0 + spos)[0] == sid
        seen.add(sid)
    assert {ID_INFO, ID_TRACKS, ID_ATTACH, ID_CUES} <= seen
    # Clusters: timestamp first, SimpleBlocks with int16 offsets, per-track times never go backwards
    cluster_time = {}
    last = {}
    nblocks = 0
    for eid, q, size, p0 in level1:
        if eid != ID_CLUSTER:
            continue
        kids = list(children(b, q, q + size))
        assert kids[0][0] == ID_TIMESTAMP
        ct = uint(b, kids[0][1], kids[0][2])
        cluster_time[p0 - seg0] = ct
        for e2, q2, s2, _ in kids[1:]:
            assert e2 == ID_SIMPLEBLOCK
            track = b[q2] & 0x7F
            rel = int.from_bytes(b[q2 + 1:q2 + 3], "big", signed=True)
            assert b[q2 + 3] & 0x80                                      # keyframe flag
            t = ct + rel
            assert t >= last.get(track, 0)
            last[track] = t
            nblocks += 1
    assert len(cluster_time) >= 2 and nblocks > 400 and set(last) == {1, 2}
    assert max(last.values()) >= 11900                                   # 12 s of audio, 1 ms time base
    # Cues: one CuePoint per Cluster, CueClusterPosition -> that Cluster, CueTime == its Timestamp
    _, q, size, _ = level1[ids.index(ID_CUES)]
    cued = set()
    for eid, q2, s2, _ in children(b, q, q + size):
        assert eid == ID_CUEPOINT
        ctime = cpos = None
        for e3, q3, s3, _ in children(b, q2, q2 + s2):
            if e3 == ID_CUETIME:
                ctime = uint(b, q3, s3)
            elif e3 == ID_CUETRACKPOS:
                for e4, q4, s4, _ in children(b, q3, q3 + s3):
                    if e4 == ID_CUECLUSTERPOS:
                        cpos = uint(b, q4, s4)
        assert cpos in cluster_time and cluster_time[cpos] == ctime
        cued.add(cpos)
    assert cued == set(cluster_time)
    # and the reference's own parser accepts the attachment as reversibility data container (Matroska.cpp:523-563)
    assert b"RAWcooked reversibility data" in b
