"""PCAMV_CONFORMANT=1: the bound host + the device with the conformance switch on (tools/reftree.py::conformance_switch,
pcamv_set_conformant), checked against the reference with the same three statements corrected
(oracle/_ref/x264_dump_conformant) - byte-identical bitstream - and the point of the mode: the payload read back from the
.264 ALONE (x264_pcamv --extract-264, host/pcamv_bitstream.c) is the message the embedder hid, the seed-1 glibc rand() & 1
stream.  Off by default; the default mode stays byte-identical to the unmodified reference (tests/test_gpu_host.py)."""
import ctypes
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import refrun

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")

CASES = [
    ("cif_hex5", 352, 288, 10, 32, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", "0.2", {}),
    # skip-heavy: macroblocks pass 2 forces to P_SKIP (quirk q2 in the default mode), through the device-built reference frames
    ("cif_qp34_skips", 352, 288, 8, 2, "--qp 34 --ref 2 --keyint 250 --me dia --subme 4", "0.2", {"PCAMV_CHECK_RECON": "1"}),
    # sub-8x8 partitions: forced P_8x8 whose pass-2 analysis chose 16x8 / 8x16; CAVLC
    ("cif_p4x4_cavlc", 352, 288, 8, 32, "--qp 22 --ref 1 --keyint 250 --me hex --subme 5 --partitions all --no-cabac", "0.1", {}),
    # the host walks pass 1 itself and cross-checks the embed stage the device built (straight vector copy on both sides)
    ("cif_umh5_host_pass1", 352, 288, 8, 32, "--qp 26 --ref 3 --keyint 250 --me umh --subme 5", "0.2", {"PCAMV_HOST_PASS1": "1", "PCAMV_CHECK_EMBED": "1"}),
]


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def glibc_rand_bits(n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return np.array([libc.rand() & 1 for _ in range(n)], dtype=np.uint8)


@pytest.mark.skipif(not refrun.have_ref("x264_dump_conformant"), reason="oracle/_ref/x264_dump_conformant not built")
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conformant_mode_bitstream_and_payload_from_the_stream_alone(pcamv, cuda_lib, case, tmp_path):
    name, w, h, frames, noise, args, emrate, env = case
    wd = str(tmp_path)
    args = args + " --emrate " + emrate
    clip = refrun.synth_clip(pcamv, w, h, frames, config=1, stream=1, noise16=noise, workdir=wd)
    ref_out, _ = refrun.run_ref(clip, w, h, args.split(), binary="x264_dump_conformant", out=os.path.join(wd, name + "_ref.264"))
    out, stats = os.path.join(wd, name + "_gpu.264"), os.path.join(wd, name + "_stats.json")
    p = subprocess.run([HOST] + args.split() + ["-o", out, clip, "%dx%d" % (w, h)], capture_output=True, timeout=1800,
                       env=dict(os.environ, PCAMV_CONFORMANT="1", PCAMV_STATS=stats, **env))
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    assert md5(out) == md5(ref_out), "bitstream differs from the conformant reference (%d vs %d bytes)" % (os.path.getsize(out), os.path.getsize(ref_out))
    st = json.load(open(stats))
    assert st["gpu_launches"] > 0 and st.get("recon_mismatch", 0) == 0 and st.get("stale_mismatch", 0) == 0, st
    # the decoder side: nothing but the stream and the rate
    msg = os.path.join(wd, "message.bin")
    p = subprocess.run([HOST, "--extract-264", out, "--emrate", emrate, "-o", msg], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:].decode("latin-1")
    raw, pos, bits, n_frames = open(msg, "rb").read(), 0, [], 0
    while pos < len(raw):
        _, an = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=2, offset=pos)]
        bits.append(np.frombuffer(raw, dtype=np.uint8, count=an, offset=pos + 8))
        pos += 8 + an
        n_frames += 1
    payload = np.concatenate(bits)
    assert n_frames == frames - 1 and len(payload) > 100
    assert np.array_equal(payload, glibc_rand_bits(len(payload))), "payload read from the .264 differs from the embedded message"
    # and the switch really changes something: the default mode's stream is the unmodified reference's, not this one
    ref_default, _ = refrun.run_ref(clip, w, h, args.split(), binary="x264_wide", out=os.path.join(wd, name + "_default.264"))
    assert md5(ref_default) != md5(ref_out)
