#!/usr/bin/env python3
"""BASELINE config 5 as it is worded - "embed + extract round trip" - with nothing but the .264 on the extraction side.

    python tools/round_trip_job.py config5            (one GPU; PCAMV_CONFORMANT=1 is set for the encoder)
    python tools/round_trip_job.py tiny --encoder reference       (CPU: the reference with the conformance switch compiled in,
                                                                    oracle/_ref/x264_dump_conformant - the test of this tool)

1. every stream / GOP of the job (video-steganography-pcamv_b200/encjob.py::JOBS) is encoded + embedded by host/_build/x264_pcamv
   in conformant mode (the default, byte-identical mode inherits streams no decoder reads back to the embedder's
   vectors from the reference);
2. every shard's NAL stream goes through `x264_pcamv --extract-264 --emrate R` (host/pcamv_bitstream.c) on the host cores;
3. the payload read back must be the message the embedder hid - the seed-1 glibc rand() & 1 stream, restarted per shard
   (encoder/encoder.c:1838-1840) - for EVERY P frame, and equal to the encoder-side record of the same run (PCAMV_PAYLOAD).
One JSON line: frames/s of the encode (wall clock of the encoder process), frames/s of the extraction, bits, verdict."""
import ctypes
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def glibc_rand_bits(n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return np.array([libc.rand() & 1 for _ in range(n)], dtype=np.uint8)


def read_messages(path):
    raw, pos, out = open(path, "rb").read(), 0, []
    while pos < len(raw):
        frame, an = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=2, offset=pos)]
        out.append((frame, an, np.frombuffer(raw, dtype=np.uint8, count=an, offset=pos + 8)))
        pos += 8 + an
    return out


def main():
    import pcamv_loader
    pcamv = pcamv_loader.load()
    from pcamv_b200 import encjob
    name = sys.argv[1]
    encoder = sys.argv[sys.argv.index("--encoder") + 1] if "--encoder" in sys.argv else "gpu"
    job = encjob.JOBS[name]
    w, h, k, n = job["width"], job["height"], job["shard_frames"], job["shards"]
    emrate = job["args"].split()[job["args"].split().index("--emrate") + 1]
    t0 = time.perf_counter()
    clip = encjob.make_clip(pcamv, name)
    t_clip = time.perf_counter() - t0
    d = encjob.job_dir(name)
    if encoder == "gpu":
        os.environ["PCAMV_CONFORMANT"] = "1"
        secs, recs, err = encjob.run_rank(name, 0, 1, int(os.environ.get("PCAMV_DEVICE", "0")), tag="roundtrip")
        files = [r["file"] for r in recs]
        side = [np.frombuffer(r["payload"], dtype=np.uint8) for r in recs]
        loop = max((r["stats"].get("t_total", 0) - r["stats"].get("t_open", 0)) for r in recs) if recs else 0
    else:
        exe = os.path.join(ROOT, "oracle", "_ref", "x264_dump_conformant")
        files = [os.path.join(d, "roundtrip_ref.264.%d" % g) for g in range(n)]

        def enc(g):
            subprocess.run([exe] + encjob.job_args(job) + ["--seek", str(g * k), "--frames", str(k), "-o", files[g], clip, "%dx%d" % (w, h)],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max(1, min(os.cpu_count() or 1, 32))) as ex:
            list(ex.map(enc, range(n)))
        secs, side, loop = time.perf_counter() - t0, None, 0

    def extract(g):
        out = files[g] + ".message"
        p = subprocess.run([encjob.HOST, "--extract-264", files[g], "--emrate", emrate, "-o", out], capture_output=True)
        if p.returncode:
            raise RuntimeError("--extract-264 failed on shard %d: %s" % (g, p.stderr[-400:].decode("latin-1")))
        return read_messages(out)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max(1, min(os.cpu_count() or 1, 32))) as ex:
        msgs = list(ex.map(extract, range(n)))
    t_extract = time.perf_counter() - t0
    bits = frames = bad_shards = skipped = side_mismatch = 0
    for g, m in enumerate(msgs):
        payload = np.concatenate([x[2] for x in m]) if m else np.zeros(0, np.uint8)
        frames += len(m); bits += len(payload); skipped += sum(1 for x in m if x[1] == 0)
        ok = len(m) == k - 1 and np.array_equal(payload, glibc_rand_bits(len(payload)))
        if side is not None and not np.array_equal(payload, side[g]):
            side_mismatch += 1
        bad_shards += not ok
    res = {"job": name, "what": job["what"] + ", embed + extract round trip through the .264 alone", "encoder": "x264_pcamv, PCAMV_CONFORMANT=1" if encoder == "gpu" else "oracle/_ref/x264_dump_conformant (CPU)",
           "shards": n, "frames": n * k, "p_frames_extracted": frames, "payload_bits": bits, "frames_without_payload": skipped,
           "encode_seconds": round(secs, 3), "encode_fps_wall": round(n * k / secs, 2), "encode_fps_loop": round(n * k / loop, 2) if loop else None,
           "extract_seconds": round(t_extract, 3), "extract_fps": round(n * k / t_extract, 1), "clip_seconds": round(t_clip, 2),
           "payload_equals_embedded_message": bad_shards == 0, "shards_with_wrong_payload": bad_shards,
           "payload_equals_encoder_side_record": None if side is None else side_mismatch == 0}
    print(json.dumps(res))
    return 0 if bad_shards == 0 and side_mismatch == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
