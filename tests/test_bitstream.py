"""The decoder side of the payload channel (host/pcamv_bitstream.c): `x264_pcamv --parse-mv` and `--extract-264`.

SURVEY.md 8(f) row 4: a bitstream-side motion-vector parser (the CABAC / CAVLC macroblock layer the reference WRITES, read back)
and extraction of the payload from the .264 alone.  The reference has no decoder, so the parser is pinned against the reference
ENCODER: the vectors it reads out of a stream must be the vectors the encoder's own analysis recorded for that stream
('MBAN' records of oracle/_ref/x264_dump, the reference compiled with dump hooks).

Three facts are pinned here, all measured on the reference itself:

1. Without embedding the reference writes conformant streams and the parser returns every macroblock's type, partitioning,
   references and vectors exactly - CABAC and CAVLC, 1-4 references, sub-8x8 partitions, QP 1 - 51, IDR GOPs.
2. WITH embedding, pass 2 of the reference (the pass whose bits are kept) leaves encoder state behind that no decoder can
   know - h->mb.i_partition of a macroblock forced to P_8x8 (encoder/analyse.c:2871-2875 sets it for P_L0 only, and
   x264_mb_predict_mv, common/macroblock.c:51-79, then applies the 16x8 / 8x16 prediction shortcuts to 8x8 blocks), the
   vector cache of a macroblock forced to P_SKIP (analyse.c:2677-2680, quirk q2) - and its cover bit comes from another
   partition's vector.  A standard decoder reads DIFFERENT vectors than the embedder meant, and the payload
   of most frames cannot be recovered from the reference's own output.  This is a property of the reference (and therefore of
   the byte-identical GPU encoder in its default mode); the test records it so that nobody mistakes it for a parser fault.
3. With those three statements corrected (tools/reftree.py::conformance_switch; oracle/_ref/x264_dump_conformant here,
   PCAMV_CONFORMANT=1 in the bound host) the parser again returns the encoder's vectors exactly, the stego vector read from the
   .264 is the embedder's, and the extracted payload is the embedded message: the seed-1 glibc rand() & 1 stream."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")

pytestmark = [pytest.mark.skipif(not os.path.exists(HOST), reason="host/_build/x264_pcamv not built (host/build_host.py)"),
              pytest.mark.skipif(not refrun.have_ref("x264_dump_conformant"), reason="oracle/_ref not built (needs /root/reference)")]

MVREC = np.dtype([("type", "<i4"), ("partition", "<i4"), ("sub", "u1", 4), ("ref", "i1", 4), ("mv", "<i2", (16, 2))])
assert MVREC.itemsize == 80
P_L0, P_8x8, P_SKIP = 4, 5, 6


def parse_mv(stream, workdir):
    out = os.path.join(workdir, "mv.bin")
    p = subprocess.run([HOST, "--parse-mv", stream, "-o", out], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    raw, pos, pics = open(out, "rb").read(), 0, []
    while pos < len(raw):
        hd = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=6, offset=pos)]
        pos += 24
        pics.append({"picture": hd[0], "n_mb": hd[1], "mb_w": hd[2], "mb_h": hd[3], "qp": hd[4], "cabac": hd[5],
                     "mb": np.frombuffer(raw, dtype=MVREC, count=hd[1], offset=pos)})
        pos += 80 * hd[1]
    return pics


def extract_264(stream, emrate, workdir):
    msg, stego = os.path.join(workdir, "message.bin"), os.path.join(workdir, "stego.bin")
    p = subprocess.run([HOST, "--extract-264", stream, "--emrate", emrate, "--stego", stego, "-o", msg], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    raw, pos, m = open(msg, "rb").read(), 0, []
    while pos < len(raw):
        frame, an = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=2, offset=pos)]
        pos += 8
        m.append((frame, an, np.frombuffer(raw, dtype=np.uint8, count=an, offset=pos)))
        pos += an
    raw, pos, s = open(stego, "rb").read(), 0, []
    while pos < len(raw):
        frame, n, an = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=3, offset=pos)]
        pos += 12
        s.append((frame, n, an, np.frombuffer(raw, dtype=np.uint8, count=n, offset=pos)))
        pos += n
    return m, s


def encode(pcamv, binary, size, frames, args, noise, stream, workdir):
    w, h = size
    clip = refrun.synth_clip(pcamv, w, h, frames, config=1, stream=stream, noise16=noise, workdir=workdir)
    dump = os.path.join(workdir, "dump.bin")
    out, _ = refrun.run_ref(clip, w, h, args.split(), binary=binary, dump=dump, planes=False, calls=False)
    return out, pcamv.dumpfmt.Dump(dump)


def count_vector_mismatches(pics, dump):
    """Parsed pictures against the encoder's records of the pass whose bits were written (the last pass of each frame)."""
    mban = dump.mb_decisions()
    bad = total = 0
    for pic in pics:
        of_frame = mban[mban["frame"] == pic["picture"]]
        final = of_frame[of_frame["pass_"] == of_frame["pass_"].max()]
        assert len(final) == pic["n_mb"] == len(pic["mb"])
        for a, b in zip(final, pic["mb"]):
            ok = a["type"] == b["type"]
            if ok and a["type"] == P_L0:
                ok = a["partition"] == b["partition"]
            if ok and a["type"] == P_8x8:
                ok = np.array_equal(a["sub"], b["sub"])
            if ok and a["type"] == P_SKIP:
                ok = np.array_equal(a["pskip_mv"], b["mv"][0])
            elif ok:
                ok = np.array_equal(a["mv"], b["mv"]) and np.array_equal(a["ref"], b["ref"])
            bad += not ok
            total += 1
    return bad, total


# (size, frames, encoder options, noise of the synthetic clip)
SYNTAX_CASES = [
    pytest.param((176, 144), 4, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", 32, id="qcif-cabac"),
    pytest.param((176, 144), 4, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --no-cabac", 32, id="qcif-cavlc"),
    pytest.param((176, 144), 5, "--qp 26 --ref 3 --keyint 250 --me umh --subme 5", 32, id="qcif-ref3-cabac"),
    pytest.param((176, 144), 5, "--qp 26 --ref 3 --keyint 250 --me umh --subme 5 --no-cabac", 32, id="qcif-ref3-cavlc"),
    pytest.param((352, 288), 3, "--qp 20 --ref 2 --keyint 250 --me hex --subme 5 --partitions all", 32, id="cif-sub8x8-cabac"),
    pytest.param((352, 288), 3, "--qp 20 --ref 2 --keyint 250 --me hex --subme 5 --partitions all --no-cabac", 32, id="cif-sub8x8-cavlc"),
    pytest.param((352, 288), 3, "--qp 10 --ref 1 --keyint 250 --me hex --subme 3 --partitions p8x8,p4x4", 64, id="cif-qp10-cabac"),
    pytest.param((352, 288), 3, "--qp 10 --ref 1 --keyint 250 --me hex --subme 3 --partitions p8x8,p4x4 --no-cabac", 64, id="cif-qp10-cavlc"),
    pytest.param((352, 288), 3, "--qp 40 --ref 1 --keyint 250 --me dia --subme 2", 4, id="cif-qp40-skips-cabac"),
    pytest.param((352, 288), 3, "--qp 40 --ref 1 --keyint 250 --me dia --subme 2 --no-cabac", 4, id="cif-qp40-skips-cavlc"),
    pytest.param((208, 112), 6, "--qp 30 --ref 4 --keyint 3 --me hex --subme 4 --partitions all", 16, id="idr-gops-ref4-cabac"),
    pytest.param((208, 112), 6, "--qp 30 --ref 4 --keyint 3 --me hex --subme 4 --partitions all --no-cabac --nf", 16, id="idr-gops-ref4-cavlc"),
    pytest.param((360, 270), 3, "--qp 26 --ref 2 --keyint 250 --me hex --subme 5", 32, id="non-mod16-size-cabac"),
    pytest.param((640, 368), 2, "--qp 51 --ref 1 --keyint 250 --me hex --subme 5", 32, id="qp51-cabac"),
    pytest.param((640, 368), 2, "--qp 1 --ref 1 --keyint 250 --me hex --subme 5", 32, id="qp1-escapes-cabac"),
    pytest.param((640, 368), 2, "--qp 1 --ref 1 --keyint 250 --me hex --subme 5 --no-cabac", 32, id="qp1-escapes-cavlc"),
]


@pytest.mark.parametrize("size,frames,args,noise", SYNTAX_CASES)
def test_parser_returns_the_reference_encoders_vectors(pcamv, size, frames, args, noise, tmp_path):
    """Fact 1: the unmodified reference, no embedding (one pass per frame, nothing forced): every macroblock of every P picture."""
    stream, dump = encode(pcamv, "x264_dump", size, frames, args + " --emrate 0", noise, 11, str(tmp_path))
    pics = parse_mv(stream, str(tmp_path))
    assert len(pics) >= 1 and all(p["cabac"] == ("--no-cabac" not in args) for p in pics)
    bad, total = count_vector_mismatches(pics, dump)
    assert total == len(pics) * pics[0]["n_mb"] and bad == 0


@pytest.mark.parametrize("size,frames,args,noise", SYNTAX_CASES)
def test_parser_returns_the_conformant_encoders_vectors_with_embedding(pcamv, size, frames, args, noise, tmp_path):
    """Fact 3, vectors: both PCAMV passes, forced decisions, flipped vectors - with the conformance switch on."""
    stream, dump = encode(pcamv, "x264_dump_conformant", size, frames, args + " --emrate 0.2", noise, 12, str(tmp_path))
    pics = parse_mv(stream, str(tmp_path))
    bad, total = count_vector_mismatches(pics, dump)
    assert total > 0 and bad == 0


def glibc_rand_bits(n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return np.array([libc.rand() & 1 for _ in range(n)], dtype=np.uint8)


EXTRACT_CASES = [
    pytest.param((352, 288), 8, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", "0.2", id="cif-cabac-0.2bpmv"),
    pytest.param((352, 288), 6, "--qp 20 --ref 1 --keyint 250 --me hex --subme 5 --partitions all --no-cabac", "0.1", id="cif-sub8x8-cavlc-0.1bpmv"),
    pytest.param((352, 288), 6, "--qp 26 --ref 2 --keyint 250 --me umh --subme 5", "60", id="cif-ref2-60-bits-per-frame"),
    pytest.param((352, 288), 6, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", "0.04", id="cif-generated-matrices"),
]


@pytest.mark.parametrize("size,frames,args,emrate", EXTRACT_CASES)
def test_payload_from_the_bitstream_alone(pcamv, size, frames, args, emrate, tmp_path):
    """Fact 3, payload: .264 in, message out - nothing from the encoder side but the embedding rate (the 'key' of the scheme)."""
    stream, dump = encode(pcamv, "x264_dump_conformant", size, frames, args + " --emrate " + emrate, 32, 13, str(tmp_path))
    embeds = dump.embeds()
    messages, stegos = extract_264(stream, emrate, str(tmp_path))
    assert len(messages) == len(stegos) == len(embeds) == frames - 1
    for e, (frame, an, msg), (_, n, _, stego) in zip(embeds, messages, stegos):
        assert frame == e["frame"]
        assert n == e["length"] and np.array_equal(stego, e["stego"]), "frame %d: stego vector read from the stream differs" % frame
        assert an == e["an"] and np.array_equal(msg, e["message"][:an]), "frame %d: payload differs" % frame
    payload = np.concatenate([m for _, _, m in messages])
    assert len(payload) > 100 and np.array_equal(payload, glibc_rand_bits(len(payload)))


def test_reference_streams_do_not_decode_to_the_embedders_vectors(pcamv, tmp_path):
    """Fact 2, recorded: the unmodified reference with embedding on.  The carriers are all there (same partitioning in both
    passes), but a decoder reads other vectors than the embedder meant for some of them, and a frame's payload survives only
    when none of its stego bits is hit.  If this test ever fails because everything matches, the reference changed."""
    args = "--qp 20 --ref 1 --keyint 250 --me hex --subme 5 --partitions all"
    stream, dump = encode(pcamv, "x264_dump", (352, 288), 6, args + " --emrate 0.1", 32, 13, str(tmp_path))
    bad, total = count_vector_mismatches(parse_mv(stream, str(tmp_path)), dump)
    assert 0 < bad < total
    embeds = dump.embeds()
    messages, stegos = extract_264(stream, "0.1", str(tmp_path))
    wrong_bits = lost = 0
    for e, (_, an, msg), (_, n, _, stego) in zip(embeds, messages, stegos):
        assert n == e["length"] and an == e["an"]
        d = int(np.count_nonzero(stego != e["stego"]))
        wrong_bits += d
        lost += not np.array_equal(msg, e["message"][:an])
        if d == 0:
            assert np.array_equal(msg, e["message"][:an])
    assert wrong_bits > 0 and lost > 0


def test_parser_refuses_what_it_does_not_read(pcamv, tmp_path):
    stream, _ = encode(pcamv, "x264_dump", (176, 144), 3, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0", 32, 14, str(tmp_path))
    data = open(stream, "rb").read()
    # a stream cut inside the last picture: loud error, no output claimed
    cut = str(tmp_path / "cut.264")
    open(cut, "wb").write(data[:len(data) - 40])
    p = subprocess.run([HOST, "--parse-mv", cut, "-o", str(tmp_path / "cut.bin")], capture_output=True, timeout=60)
    assert p.returncode != 0 and b"x264 [error]" in p.stderr
    # no parameter sets: the slices cannot be read
    first_slice = data.index(b"\x00\x00\x00\x01\x65")
    headless = str(tmp_path / "headless.264")
    open(headless, "wb").write(data[first_slice:])
    p = subprocess.run([HOST, "--parse-mv", headless, "-o", str(tmp_path / "x.bin")], capture_output=True, timeout=60)
    assert p.returncode != 0 and b"missing parameter set" in p.stderr
    # --extract-264 needs the rate
    p = subprocess.run([HOST, "--extract-264", stream, "-o", str(tmp_path / "m.bin")], capture_output=True, timeout=60)
    assert p.returncode != 0 and b"--emrate" in p.stderr


@pytest.mark.skipif(not os.path.isdir(os.environ.get("PCAMV_REFERENCE", "/root/reference")) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libx264_wide.a")),
                    reason="needs the reference's headers and oracle/_ref/libx264_wide.a to build the sanitizer harness")
def test_parser_survives_hostile_streams():
    """The parser reads untrusted bytes: built with AddressSanitizer + UBSan and fed mutated streams (byte flips, truncations,
    splices, garbage behind valid headers), it must parse or refuse with a message - no sanitizer report, no signal, no leak."""
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bitstream_fuzz.py"), "120", "7"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "0 sanitizer reports, 0 signals" in p.stdout


def test_extractor_recognises_frames_the_embedder_gave_up_on(pcamv, tmp_path):
    """Where stc_embed gives up (here: messages of a few bits on a 64x48 frame) the reference leaves the stego vector zeroed and pass
    2 still flips every carrier whose cover bit is 1 - every carrier of the written picture is even.  The decoder side must report
    such a picture as carrying nothing instead of inventing a payload."""
    args = "--qp 38 --ref 3 --keyint 250 --me dia --subme 1 --no-dct-decimate"
    stream, dump = encode(pcamv, "x264_dump_conformant", (64, 48), 4, args + " --emrate 0.1", 32, 461, str(tmp_path))
    embeds = dump.embeds()
    messages, stegos = extract_264(stream, "0.1", str(tmp_path))
    gave_up = 0
    for e, (_, an, _), (_, n, _, stego) in zip(embeds, messages, stegos):
        assert n == e["length"] and np.array_equal(stego, e["stego"])
        if 1 <= e["an"] <= e["length"] and e["length"] >= 16 and not e["stego"].any() and e["cover"].any():
            gave_up += 1
            assert an == 0
    assert gave_up >= 1


def test_parse_mv_csv_is_the_same_motion_field(pcamv, tmp_path):
    """`--parse-mv --csv`: the text form (one line per macroblock) carries exactly the binary records."""
    stream, _ = encode(pcamv, "x264_dump_conformant", (176, 144), 3, "--qp 26 --ref 2 --keyint 250 --me hex --subme 5 --partitions all --emrate 0.2", 32, 15, str(tmp_path))
    pics = parse_mv(stream, str(tmp_path))
    out = str(tmp_path / "mv.csv")
    p = subprocess.run([HOST, "--parse-mv", stream, "-o", out, "--csv"], capture_output=True, timeout=60)
    assert p.returncode == 0, p.stderr[-1000:].decode("latin-1")
    rows = np.loadtxt(out, delimiter=",", dtype=np.int64)
    assert rows.shape == (sum(pc["n_mb"] for pc in pics), 13 + 32)
    k = 0
    for pc in pics:
        for mb, r in enumerate(pc["mb"]):
            row = rows[k]; k += 1
            assert row[0] == pc["picture"] and row[1] == mb % pc["mb_w"] and row[2] == mb // pc["mb_w"]
            assert row[3] == r["type"] and row[4] == r["partition"] and list(row[5:9]) == list(r["sub"]) and list(row[9:13]) == list(r["ref"])
            assert np.array_equal(row[13:].reshape(16, 2), r["mv"])
