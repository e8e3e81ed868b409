"""Whole-encoder jobs (BASELINE.json configs 2 / 4 / 5): a clip of IDR-bounded shards — GOPs of one long clip, or the
independent streams of a batch — encoded + embedded by `host/_build/x264_pcamv` (the reference's C host bound to
libpcamv_cuda.so), sharded over the ranks of one box.

  shard g = frames [g*K, (g+1)*K) of the job clip (the reference CLI's own --seek / --frames, x264.c:546-551,857); an IDR
  resets reference list, frame_num and POC (encoder/encoder.c:2247-2254), so with constant QP every shard is an independent
  encoder run.  Rank r of N runs `x264_pcamv --shards n --shard-frames K --shard-first r --shard-step N`
  with PCAMV_DEVICE=local_rank: its shards are encoder threads of one process that share the GPU through encoder groups
  (multi-context launches).  The one exchange step is `shard.gather_gop_results` (two NCCL all_gathers): per-shard payload
  bits, statistics and the md5 of the shard's NAL stream travel to rank 0 INSIDE the timed region.

  Parity: the concatenation, in shard order, of per-shard runs of the reference encoder (`oracle/_ref/x264_wide`, the
  reference's own C sources) on the same clip — bitstream md5 per shard, and the payload bits per frame against the
  instrumented twin (`x264_dump`) where asked.  The reference side is also the CPU arm (one process per host core) and is
  cached per box under /tmp (both bench arms and every N of a scaling run use the same job).

Only bench.py and tests/ use this module; nothing here is on the product path (that is the C host + the library)."""
import hashlib
import json
import os
import subprocess
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")
REF = os.path.join(ROOT, "oracle", "_ref", "x264_wide")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "x264_dump")

ENC = "--qp 26 --ref 1 --me umh --subme 5 --emrate 0.2"
# name -> geometry, shards, frames per shard, synth config, encoder arguments (the --keyint is the shard length: IDR-bounded GOPs)
JOBS = {
    # BASELINE config 2: 1080p x 300 frames, umh, single reference, 0.2 bits/MV (subme 5, not 7: RD mode decision is out of scope)
    "config2": dict(width=1920, height=1080, shards=10, shard_frames=30, synth=2, args=ENC, what="1080p x 300 frames, 10 GOPs of 30"),
    # BASELINE config 4: 3840x2160 x 600 frames, --keyint 75 -> 8 IDR-bounded GOPs
    "config4": dict(width=3840, height=2160, shards=8, shard_frames=75, synth=4, args=ENC, what="3840x2160 x 600 frames, 8 GOPs of 75"),
    "config4-k30": dict(width=3840, height=2160, shards=20, shard_frames=30, synth=4, args=ENC, what="3840x2160 x 600 frames, 20 GOPs of 30"),
    # BASELINE config 5: 64 independent 720p streams, embed + extract round trip
    "config5": dict(width=1280, height=720, shards=64, shard_frames=30, synth=5, args=ENC, streams=True, what="64 streams of 1280x720 x 30 frames"),
    # bounded jobs for the default bench run (same shapes, fewer frames)
    "config4-small": dict(width=3840, height=2160, shards=8, shard_frames=12, synth=4, args=ENC, what="3840x2160 x 96 frames, 8 GOPs of 12"),
    "config2-small": dict(width=1920, height=1080, shards=16, shard_frames=8, synth=2, args=ENC, what="1080p x 128 frames, 16 GOPs of 8"),
    "config5-small": dict(width=1280, height=720, shards=16, shard_frames=5, synth=5, args=ENC, streams=True, what="16 streams of 1280x720 x 5 frames"),
    # CPU-sized job for the tests
    "tiny": dict(width=352, height=288, shards=4, shard_frames=4, synth=1, args="--qp 26 --ref 1 --me hex --subme 5 --emrate 0.2", what="CIF x 16 frames, 4 GOPs of 4"),
}


def job_dir(name):
    d = os.path.join(os.environ.get("PCAMV_JOB_DIR", os.path.join(tempfile.gettempdir(), "pcamv_jobs")), name)
    os.makedirs(d, exist_ok=True)
    return d


def job_args(job):
    return job["args"].split() + ["--keyint", str(job["shard_frames"])]


def make_clip(pcamv, name, workers=None):
    """The job clip: shard g is generated with its own seed (stream = g) straight into its place in one file, by parallel
    generator processes.  Returns the path (generated once per box)."""
    job = JOBS[name]
    d = job_dir(name)
    path = os.path.join(d, "clip.yuv")
    done = path + ".done"
    if os.path.exists(done):
        return path
    exe = pcamv.build.build_synth()
    w, h, k, n = job["width"], job["height"], job["shard_frames"], job["shards"]
    with open(path, "wb") as f:
        f.truncate((w * h * 3 // 2) * k * n)
    workers = workers or max(1, min(os.cpu_count() or 1, 32))

    def gen(g):
        subprocess.check_call([exe, str(w), str(h), str(k), str(job["synth"]), str(g), path, "32", str(g * k)])
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(gen, range(n)))
    open(done, "w").write("ok\n")
    return path


def _md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def read_payload(path):
    """PCAMV_PAYLOAD side file -> list of (frame, length, an, message bits, stego bits)."""
    if not os.path.exists(path):
        return []
    raw = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(raw):
        frame, length, an = np.frombuffer(raw, dtype="<i4", count=3, offset=pos); pos += 12
        n = max(int(an), 0)
        msg = np.frombuffer(raw, dtype=np.uint8, count=n, offset=pos); pos += n
        stego = np.frombuffer(raw, dtype=np.uint8, count=int(length), offset=pos); pos += int(length)
        out.append((int(frame), int(length), int(an), msg, stego))
    return out


def reference_side(pcamv, name, cores=None, want_payload=True):
    """Per-shard runs of the reference encoder on all host cores (one single-threaded process per core: frame threads crash with
    embedding on).  Returns {"md5": [...], "bytes": [...], "seconds": wall, "cores": c, "payload": [...]}; cached."""
    job = JOBS[name]
    d = job_dir(name)
    cache = os.path.join(d, "reference.json")
    if os.path.exists(cache):
        return json.load(open(cache))
    # PCAMV_JOB_DIGESTS=<dir>: committed per-shard digests of the reference's output (tools/reference_digests.py; the clips are
    # seeded and the reference is deterministic) instead of running it on this box; its fps then is the fps of the machine named
    # in "made_on", and the result says so
    shipped = os.environ.get("PCAMV_JOB_DIGESTS")
    if shipped and os.path.exists(os.path.join(shipped, name + ".json")):
        res = json.load(open(os.path.join(shipped, name + ".json")))
        res["source"] = "committed digests (%s), reference not run on this box" % res.get("made_on", "?")
        make_clip(pcamv, name)
        return res
    clip = make_clip(pcamv, name)
    w, h, k, n = job["width"], job["height"], job["shard_frames"], job["shards"]
    cores = cores or max(1, min(os.cpu_count() or 1, 64))
    outs = [os.path.join(d, "ref_%d.264" % g) for g in range(n)]

    def run(g):
        subprocess.run([REF] + job_args(job) + ["--seek", str(g * k), "--frames", str(k), "-o", outs[g], clip, "%dx%d" % (w, h)],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    # the hidden payload per shard comes from the instrumented twin of the reference (same sources + dump hooks), untimed; with
    # cores to spare it runs beside the timed processes instead of after them
    def pay(g):
        dump = os.path.join(d, "ref_%d.dump" % g)
        env = dict(os.environ, PCAMV_DUMP=dump, PCAMV_DUMP_PLANES="0", PCAMV_DUMP_CALLS="0")
        subprocess.run([REF_DUMP] + job_args(job) + ["--seek", str(g * k), "--frames", str(k), "-o", os.devnull, clip, "%dx%d" % (w, h)],
                       env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        emb = pcamv.dumpfmt.Dump(dump).embeds()
        os.remove(dump)
        m = hashlib.md5()
        bits = 0
        for e in emb:
            an = max(int(e["an"]), 0)
            m.update(np.asarray(e["message"][:an], dtype=np.uint8).tobytes())
            m.update(np.asarray(e["stego"], dtype=np.uint8).tobytes())
            bits += an
        return m.hexdigest(), bits
    side = None
    if want_payload and (os.cpu_count() or 1) >= 2 * min(cores, n):
        side = ThreadPoolExecutor(min(cores, n))
        side_futs = [side.submit(pay, g) for g in range(n)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(run, range(n)))
    wall = time.perf_counter() - t0
    res = {"md5": [_md5(o) for o in outs], "bytes": [os.path.getsize(o) for o in outs], "seconds": wall, "cores": min(cores, n),
           "frames": n * k, "fps": n * k / wall, "payload_md5": None,
           "concurrent_payload_runs": side is not None}
    if want_payload:
        if side is not None:
            pm = [f.result() for f in side_futs]
            side.shutdown()
        else:
            with ThreadPoolExecutor(cores) as ex:
                pm = list(ex.map(pay, range(n)))
        res["payload_md5"] = [p[0] for p in pm]
        res["payload_bits"] = [p[1] for p in pm]
    for o in outs:
        os.remove(o)
    json.dump(res, open(cache, "w"))
    return res


def run_rank(name, rank, world, device, groups=4, extract=False, tag="gpu"):
    """This rank's share of the job through x264_pcamv.  Returns (seconds of the encoder process, per-shard records)."""
    job = JOBS[name]
    d = job_dir(name)
    clip = os.path.join(d, "clip.yuv")
    w, h, k, n = job["width"], job["height"], job["shard_frames"], job["shards"]
    mine = list(range(rank, n, world))
    if not mine:
        return 0.0, [], ""
    out = os.path.join(d, "%s_r%d.264" % (tag, rank))
    pay = os.path.join(d, "%s_payload_r%d" % (tag, rank))
    for g in mine:
        for f in ("%s.%d" % (out, g), "%s.%d" % (pay, g)):
            if os.path.exists(f):
                os.remove(f)
    st = os.path.join(d, "%s_stats_r%d" % (tag, rank))
    groups = int(os.environ.get("PCAMV_JOB_GROUPS", groups))           # (tuning: rendezvous groups and wavefront layout of the job's encoders)
    env = dict(os.environ, PCAMV_DEVICE=str(device), PCAMV_ROWS_PER_CTA=os.environ.get("PCAMV_JOB_RPC", "4"), PCAMV_GROUPS=str(max(1, min(groups, len(mine)))),
               PCAMV_PAYLOAD=pay, PCAMV_STATS=st)
    env.pop("CUDA_DEVICE_MAX_CONNECTIONS", None)
    cmd = [HOST, "--shards", str(len(mine)), "--shard-frames", str(k), "--shard-first", str(rank), "--shard-step", str(world), "--shard-keep"] + \
        job_args(job) + ["-o", out, clip, "%dx%d" % (w, h)]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, env=env, capture_output=True)
    secs = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("x264_pcamv failed on rank %d: %s" % (rank, p.stderr[-600:].decode("latin-1")))
    recs = []
    for g in mine:
        f = "%s.%d" % (out, g)
        pl = read_payload("%s.%d" % (pay, g))
        m = hashlib.md5()
        bits, n_mv, payload = 0, 0, []
        for (_, length, an, msg, stego) in pl:
            m.update(msg.tobytes()); m.update(stego.tobytes())
            bits += max(an, 0); n_mv += length
            payload.append(msg)
        try:
            stats = json.load(open("%s.%d" % (st, g)))
        except Exception:
            stats = {}
        recs.append({"stats": stats, "gop": g, "n_bits": bits, "payload": np.concatenate(payload).tobytes() if payload else b"", "n_mv": n_mv,
                     "n_flipped": 0, "bytes": os.path.getsize(f), "md5": _md5(f), "payload_md5": m.hexdigest(), "file": f})
    return secs, recs, p.stderr.decode("latin-1")
