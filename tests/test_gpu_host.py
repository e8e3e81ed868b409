"""Whole-encoder parity: the reference's C host with the CUDA shim bound in (host/_build/x264_pcamv, built by
host/build_host.py) must emit a .264 bitstream that is byte-identical to the reference encoder's (oracle/_ref/x264_ref
for CIF, the widened build otherwise) on the same synthetic clip and flags — both PCAMV passes, embedding on."""
import hashlib
import json
import os
import subprocess

import pytest

import refrun

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")

CASES = [
    # BASELINE.json config 1: CIF 352x288, 30 frames, --me hex --subme 5, embed at 0.2 bits/MV; the UNMODIFIED reference
    ("cif_hex5", 352, 288, 30, 1, 32, "x264_ref", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2"),
    ("cif_umh5_ref3", 352, 288, 12, 1, 32, "x264_wide", "--qp 26 --ref 3 --keyint 250 --me umh --subme 5 --emrate 0.2"),
    ("cif_dia2_lownoise", 352, 288, 12, 1, 4, "x264_wide", "--qp 32 --ref 1 --keyint 250 --me dia --subme 2 --emrate 0.2"),
    ("cif_hex4_idr", 352, 288, 14, 1, 16, "x264_wide", "--qp 24 --ref 2 --keyint 6 --min-keyint 6 --me hex --subme 4 --emrate 0.3"),
    ("cif_noembed", 352, 288, 8, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5"),
    ("qcif_esa", 176, 144, 6, 9, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me esa --merange 16 --subme 5 --emrate 0.2"),
    ("qcif_tesa", 176, 144, 6, 9, 32, "x264_wide", "--qp 26 --ref 2 --keyint 250 --me tesa --merange 16 --subme 5 --emrate 0.2"),
    ("cif_esa32_ref4", 352, 288, 6, 1, 32, "x264_wide", "--qp 26 --ref 4 --keyint 250 --me esa --merange 32 --subme 5 --emrate 0.2"),
    ("cif_p4x4_hex5", 352, 288, 8, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --partitions p8x8,p4x4 --emrate 0.2"),
    ("cif_p4x4_umh_ref3", 352, 288, 8, 1, 24, "x264_wide", "--qp 24 --ref 3 --keyint 250 --me umh --subme 4 --partitions all --emrate 0.3"),
    # skip-heavy (38 % P_SKIP): pass-2 probes succeed on macroblocks pass 1 coded, the host keeps b_skip_mc set (quirk q1) and the
    # residual is taken against what its intra analysis left in fdec — those macroblocks are exempt from pass-2 elision
    ("cif_qp48_skips", 352, 288, 6, 1, 24, "x264_wide", "--qp 48 --ref 2 --keyint 250 --me umh --subme 4 --emrate 0.2"),
    # tiny frames: one macroblock, one macroblock row / column, and a 3x2 frame whose messages are shorter than the STC matrix height
    ("tiny_16x16", 16, 16, 5, 1, 32, "x264_wide", "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"),
    ("tiny_1024x16", 1024, 16, 4, 1, 32, "x264_wide", "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"),
    ("tiny_16x512", 16, 512, 4, 1, 32, "x264_wide", "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"),
    ("tiny_48x32", 48, 32, 6, 1, 32, "x264_wide", "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"),
    ("720p_umh5", 1280, 720, 4, 5, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"),
    # option coverage (the reference Makefile's OPT0..OPT7 flag sets, Makefile:108-115, restricted to the supported path)
    ("cif_nocabac", 352, 288, 8, 1, 16, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --no-cabac --emrate 0.2"),
    ("cif_nofastpskip_ref4", 352, 288, 9, 1, 8, "x264_wide", "--qp 28 --ref 4 --keyint 250 --me umh --subme 4 --no-fast-pskip --emrate 0.2"),
    ("cif_subme1_dia", 352, 288, 8, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me dia --subme 1 --emrate 0.2"),
    ("cif_subme3_merange8", 352, 288, 8, 1, 32, "x264_wide", "--qp 22 --ref 2 --keyint 250 --me hex --merange 8 --subme 3 --emrate 0.5"),
    ("cif_fixed_bits", 352, 288, 8, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 40"),
    ("cif_p16x16_only", 352, 288, 8, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --partitions none --emrate 0.2"),
    ("odd_size_umh", 360, 270, 6, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"),
    ("cif_nodecimate_qp36", 352, 288, 8, 1, 8, "x264_wide", "--qp 36 --ref 1 --keyint 250 --me hex --subme 5 --no-dct-decimate --emrate 0.2"),
    # BASELINE.json config 2 / config 4 geometries (1080p is padded to 1088 lines; 4K = 240 x 135 macroblocks)
    ("1080p_umh5", 1920, 1080, 3, 2, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"),
    ("4k_hex5_ref2", 3840, 2160, 3, 4, 32, "x264_wide", "--qp 28 --ref 2 --keyint 250 --me hex --subme 5 --emrate 0.2"),
]


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def encode_pair(pcamv, name, w, h, frames, config, noise, ref_bin, args, workdir, extra_env=None):
    clip = refrun.synth_clip(pcamv, w, h, frames, config=config, stream=1, noise16=noise, workdir=workdir)
    ref_out, _ = refrun.run_ref(clip, w, h, args.split(), binary=ref_bin, out=os.path.join(workdir, name + "_ref.264"))
    out = os.path.join(workdir, name + "_gpu.264")
    stats = os.path.join(workdir, name + "_stats.json")
    env = dict(os.environ, PCAMV_STATS=stats, **(extra_env or {}))
    p = subprocess.run([HOST] + args.split() + ["-o", out, clip, "%dx%d" % (w, h)], env=env, capture_output=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")      # (the reference prints GB18030 text)
    return ref_out, out, json.load(open(stats))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_bitstream_identical(pcamv, cuda_lib, case, tmp_path):
    if not os.path.exists(HOST):
        pytest.fail("host/_build/x264_pcamv is not built (host/build_host.py)")
    ref_out, out, stats = encode_pair(pcamv, *case, workdir=str(tmp_path))
    assert os.path.getsize(out) > (1000 if case[1] * case[2] > 20000 else 100)
    assert md5(out) == md5(ref_out), "bitstream differs from the reference (%d vs %d bytes)" % (os.path.getsize(out), os.path.getsize(ref_out))
    assert stats["gpu_launches"] > 0 and stats["replayed_calls"] > 0
    if "--emrate" in case[7]:
        # pass 1 of every embedding P frame stayed on the device: the host replayed one pass per frame, not two
        assert stats["direct_pass1"] > 0 and stats["p_passes"] == 2 * stats["direct_pass1"], stats


HOST_PASS1 = [c for c in CASES if c[0] in ("cif_hex5", "cif_umh5_ref3", "cif_dia2_lownoise", "cif_p4x4_umh_ref3", "cif_qp48_skips", "cif_nocabac",
                                           "tiny_48x32", "cif_fixed_bits", "720p_umh5")]


@pytest.mark.parametrize("case", HOST_PASS1, ids=[c[0] for c in HOST_PASS1])
def test_bitstream_identical_with_host_pass1(pcamv, cuda_lib, case, tmp_path):
    """PCAMV_HOST_PASS1=1: the host walks the macroblocks of pass 1 as the reference does (replaying the GPU's results), and the
    embed stage built on the device is checked against the host's own cover / rho_final (PCAMV_CHECK_EMBED=1).  Also checks the one
    piece of host state that pass 1 hands to pass 2 outside h->info — the MV cache of its last macroblock (quirk q2) — against
    what the device-only pass 1 would have put there."""
    ref_out, out, stats = encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env={"PCAMV_HOST_PASS1": "1", "PCAMV_CHECK_EMBED": "1"})
    assert md5(out) == md5(ref_out)
    assert stats["direct_pass1"] == 0 and stats["stale_mismatch"] == 0, stats


@pytest.mark.parametrize("case", [CASES[1], CASES[2]], ids=[CASES[1][0], CASES[2][0]])
def test_bitstream_identical_with_full_pass2(pcamv, cuda_lib, case, tmp_path):
    """Same, with the GPU executing and logging every pass-2 search the reference executes (PCAMV_PASS2_FULL=1) instead of
    eliding the ones whose results the reference overwrites."""
    ref_out, out, stats = encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env={"PCAMV_PASS2_FULL": "1"})
    assert md5(out) == md5(ref_out)


@pytest.mark.parametrize("env", [{"PCAMV_NO_ROW_STREAM": "1"}, {"PCAMV_HOST_INTRA": "1"}, {"PCAMV_NO_PINNED": "1"},
                                 {"PCAMV_ROWS_PER_CTA": "4"}, {"PCAMV_BLOCKING_SYNC": "1"}, {"PCAMV_DEVICE_RECON": "0"}],
                         ids=["no-row-stream", "host-intra", "no-pinned", "row-groups-4", "blocking-sync", "uploaded-references"])
@pytest.mark.parametrize("case", [CASES[1], CASES[10]], ids=[CASES[1][0], CASES[10][0]])
def test_bitstream_identical_with_host_switches(pcamv, cuda_lib, case, env, tmp_path):
    """The switches of the bound host that change HOW it gets its results, never WHAT they are: waiting for the end of the replayed
    pass instead of following the wavefront row by row, intra analysis of every P macroblock instead of the q1 ones only, pageable
    result buffers (which also disables the row streaming), the throughput layout of the wavefront, sleeping instead of spinning
    waits — on a 3-reference umh clip and on the skip-heavy clip that exercises quirks q1 / q2."""
    ref_out, out, stats = encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env=env)
    assert md5(out) == md5(ref_out)
    if "PCAMV_NO_ROW_STREAM" in env or "PCAMV_NO_PINNED" in env:
        assert stats["t_row_wait"] == 0.0
    # by default the P frames become references on the device (and hand their half-pel planes back); the switch uploads them
    assert (stats["recon_frames"] == 0) == ("PCAMV_DEVICE_RECON" in env)


def test_payload_identical(pcamv, cuda_lib, tmp_path):
    """The hidden payload: per P frame the message bits (glibc rand() & 1 stream, encoder/encoder.c:1838-1840) and the stego
    LSB vector the STC embedder produced are identical to the reference's (EMBD records of the instrumented twin)."""
    import numpy as np
    w, h, frames, args = 352, 288, 10, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2"
    workdir = str(tmp_path)
    clip = refrun.synth_clip(pcamv, w, h, frames, config=1, stream=2, workdir=workdir)
    dump = os.path.join(workdir, "d.bin")
    refrun.run_ref(clip, w, h, args.split(), dump=dump, planes=False, calls=False)
    ref = pcamv.dumpfmt.Dump(dump).embeds()
    pay = os.path.join(workdir, "payload.bin")
    p = subprocess.run([HOST] + args.split() + ["-o", os.path.join(workdir, "o.264"), clip, "%dx%d" % (w, h)],
                       env=dict(os.environ, PCAMV_PAYLOAD=pay), capture_output=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    raw = open(pay, "rb").read()
    pos, got = 0, []
    while pos < len(raw):
        frame, length, an = np.frombuffer(raw, dtype="<i4", count=3, offset=pos); pos += 12
        msg = np.frombuffer(raw, dtype=np.uint8, count=max(int(an), 0), offset=pos); pos += max(int(an), 0)
        stego = np.frombuffer(raw, dtype=np.uint8, count=int(length), offset=pos); pos += int(length)
        got.append((int(frame), int(length), int(an), msg, stego))
    assert len(got) == len(ref) and len(got) >= frames - 1
    bits = 0
    for (frame, length, an, msg, stego), e in zip(got, ref):
        assert (frame, length, an) == (e["frame"], e["length"], e["an"])
        assert np.array_equal(msg, e["message"]) and np.array_equal(stego, e["stego"])
        bits += max(an, 0)
    assert bits > 100


def test_sharded_encode_equals_per_gop_reference_runs(pcamv, cuda_lib, tmp_path):
    """GOP sharding: `x264_pcamv --shards N --shard-frames K` runs N encoder instances on threads of one
    process, sharing the GPU through an encoder group (one multi-context launch per step); its output must be the
    concatenation of N independent reference runs `--seek g*K --frames K` (the parity definition for sharded mode)."""
    w, h, n, k = 352, 288, 4, 5
    args = "--qp 26 --ref 2 --keyint 250 --me hex --subme 5 --emrate 0.2"
    workdir = str(tmp_path)
    clip = refrun.synth_clip(pcamv, w, h, n * k, config=1, stream=4, workdir=workdir)
    want = b""
    for g in range(n):
        out, _ = refrun.run_ref(clip, w, h, args.split() + ["--seek", str(g * k), "--frames", str(k)], binary="x264_wide",
                                out=os.path.join(workdir, "ref_%d.264" % g))
        want += open(out, "rb").read()
    out = os.path.join(workdir, "sharded.264")
    p = subprocess.run([HOST, "--shards", str(n), "--shard-frames", str(k)] + args.split() + ["-o", out, clip, "%dx%d" % (w, h)],
                       capture_output=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    got = open(out, "rb").read()
    assert len(got) == len(want) and got == want


def test_static_clip_no_carriers(pcamv, cuda_lib, tmp_path):
    """Edge case: identical frames -> (almost) every macroblock is P_SKIP, frames carry few or no motion vectors, and the
    embed stage runs through its degenerate paths (an = 0, embed.h:347; stc_embed's return value is ignored,
    encoder/encoder.c:1843).  The GPU-backed encoder must still be byte-identical."""
    w, h, n = 352, 288, 6
    args = "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2"
    workdir = str(tmp_path)
    one = refrun.synth_clip(pcamv, w, h, 1, config=1, stream=7, workdir=workdir)
    clip = os.path.join(workdir, "static.yuv")
    with open(clip, "wb") as f:
        f.write(open(one, "rb").read() * n)
    ref_out, _ = refrun.run_ref(clip, w, h, args.split(), binary="x264_wide", out=os.path.join(workdir, "ref.264"))
    out = os.path.join(workdir, "gpu.264")
    p = subprocess.run([HOST] + args.split() + ["-o", out, clip, "%dx%d" % (w, h)], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    assert md5(out) == md5(ref_out)


def test_multi_stream_embed_extract_round_trip(pcamv, cuda_lib, tmp_path):
    """BASELINE config 5 in small: a batch of independent streams (different content, concatenated in one clip and encoded as
    shards of one process sharing the GPU) — every stream's bitstream equals the reference encoder's run on that stream,
    and the payload extracted from the stego vectors our encoder wrote (x264_pcamv --extract, host/pcamv_stc_extract.c)
    equals, bit for bit, the message the REFERENCE embedded in that stream."""
    import numpy as np
    import test_extract
    w, h, n, k = 640, 368, 8, 4
    args = "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
    workdir = str(tmp_path)
    clip = os.path.join(workdir, "streams.yuv")
    with open(clip, "wb") as f:
        for g in range(n):          # stream g: its own seed
            f.write(open(refrun.synth_clip(pcamv, w, h, k, config=5, stream=10 + g, workdir=workdir), "rb").read())
    want, ref_msgs = b"", []
    for g in range(n):
        dump = os.path.join(workdir, "d%d.bin" % g)
        out, _ = refrun.run_ref(clip, w, h, args.split() + ["--seek", str(g * k), "--frames", str(k)], dump=dump, planes=False,
                                calls=False, out=os.path.join(workdir, "ref_%d.264" % g))
        want += open(out, "rb").read()
        ref_msgs.append(pcamv.dumpfmt.Dump(dump).embeds())
    out, stego = os.path.join(workdir, "batch.264"), os.path.join(workdir, "stego")
    p = subprocess.run([HOST, "--shards", str(n), "--shard-frames", str(k)] + args.split() + ["-o", out, clip, "%dx%d" % (w, h)],
                       env=dict(os.environ, PCAMV_STEGO=stego), capture_output=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    assert open(out, "rb").read() == want, "batch of streams differs from the per-stream reference runs"
    bits = 0
    for g in range(n):
        msg = os.path.join(workdir, "msg_%d.bin" % g)
        q = subprocess.run([HOST, "--extract", "%s.%d" % (stego, g), "-o", msg], capture_output=True, timeout=600)
        assert q.returncode == 0, q.stderr[-2000:].decode("latin-1")
        got = test_extract.read_messages(msg)
        assert len(got) == len(ref_msgs[g]) == k - 1
        for (frame, an, m), e in zip(got, ref_msgs[g]):
            assert an == e["an"] and np.array_equal(m, e["message"][:an]), "stream %d frame %d: extracted payload differs" % (g, frame)
            bits += an
    assert bits > 1000
