"""The extraction side of the payload channel (host/pcamv_stc_extract.c, `x264_pcamv --extract`).

The reference embeds (stc_embed, embed.h:309) but ships no extractor, so SURVEY.md 8(c) defines extraction as the STC
syndrome of the stego LSB vector under the embedder's own sub-matrices.  Pinned here against the reference itself: from
the stego vectors the reference encoder produced (EMBD records of oracle/_ref/x264_dump: golden fixtures and live runs)
the extractor must return exactly the message the reference embedded — which in turn is the seed-1 glibc rand() & 1
stream (encoder/encoder.c:1838-1840), consumed frame after frame."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")

pytestmark = pytest.mark.skipif(not os.path.exists(HOST), reason="host/_build/x264_pcamv not built (host/build_host.py)")


def extract(embeds, workdir):
    """embeds: [(frame, length, an, stego uint8[length])] -> [(frame, an, message uint8[an])] through the CLI."""
    src, dst = os.path.join(workdir, "stego.bin"), os.path.join(workdir, "message.bin")
    with open(src, "wb") as f:
        for frame, length, an, stego in embeds:
            f.write(np.array([frame, length, an], dtype="<i4").tobytes())
            f.write(np.ascontiguousarray(stego, dtype=np.uint8).tobytes())
    p = subprocess.run([HOST, "--extract", src, "-o", dst], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    return read_messages(dst)


def read_messages(path):
    raw, pos, out = open(path, "rb").read(), 0, []
    while pos < len(raw):
        frame, an = np.frombuffer(raw, dtype="<i4", count=2, offset=pos); pos += 8
        out.append((int(frame), int(an), np.frombuffer(raw, dtype=np.uint8, count=int(an), offset=pos))); pos += int(an)
    return out


def glibc_rand_bits(n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return np.array([libc.rand() & 1 for _ in range(n)], dtype=np.uint8)


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_tesa5", "qcif_dia2_lownoise"])
def test_extract_recovers_reference_message_golden(pcamv, name, tmp_path):
    ref = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path))).embeds()
    got = extract([(e["frame"], e["length"], e["an"], e["stego"]) for e in ref], str(tmp_path))
    assert len(got) == len(ref) >= 1
    bits = 0
    for (frame, an, msg), e in zip(got, ref):
        assert frame == e["frame"] and an == max(e["an"], 0)
        assert np.array_equal(msg, e["message"][:an]), "frame %d: extracted payload differs from the embedded one" % frame
        bits += an
    assert bits > 20


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("emrate", ["0.2", "0.04", "60"])      # bits per MV (widths 5/6: built-in matrices; 25: generated ones), fixed bits per frame
def test_extract_round_trip_live_reference(pcamv, emrate, tmp_path):
    w, h, frames = 352, 288, 8
    clip = refrun.synth_clip(pcamv, w, h, frames, config=1, stream=7, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, w, h, ("--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate " + emrate).split(), dump=dump,
                   planes=False, calls=False)
    ref = pcamv.dumpfmt.Dump(dump).embeds()
    got = extract([(e["frame"], e["length"], e["an"], e["stego"]) for e in ref], str(tmp_path))
    assert len(got) == len(ref) == frames - 1
    payload = np.concatenate([m for _, _, m in got])
    for (frame, an, msg), e in zip(got, ref):
        assert an == e["an"] and np.array_equal(msg, e["message"][:an])
    # the whole run's payload is the seed-1 rand() & 1 stream, consumed in coding order of the P frames
    assert np.array_equal(payload, glibc_rand_bits(len(payload)))


def test_extract_rejects_truncated_input(tmp_path):
    src = str(tmp_path / "bad.bin")
    with open(src, "wb") as f:
        f.write(np.array([1, 100, 20], dtype="<i4").tobytes())
        f.write(bytes(10))
    p = subprocess.run([HOST, "--extract", src, "-o", str(tmp_path / "o.bin")], capture_output=True)
    assert p.returncode != 0 and b"truncated" in p.stderr


def test_extract_passes_unextractable_frames_by(tmp_path):
    """A frame whose header claims more message bits than it has carriers (an embedding the encoder could not have made)
    does not take the rest of the stream with it: it comes out with an = 0 and the following frames are extracted."""
    good = (2, 40, 8, np.tile(np.array([1, 0, 0, 1, 0], dtype=np.uint8), 8))
    got = extract([(1, 5, 9, np.ones(5, dtype=np.uint8)), good], str(tmp_path))
    assert [(f, an) for f, an, _ in got] == [(1, 0), (2, 8)]
    alone = extract([good], str(tmp_path))
    assert np.array_equal(got[1][2], alone[0][2])
